// adapter_smoke.cpp -- drives vloam::LidarOdometryMapping (vloam_adapter.hpp) exactly as
// vloam_main_node.cpp:143-144, 186-190 drives the reference, on sweeps read from a raw float file.
// usage: adapter_smoke <scan0.bin> <scan1.bin>   (float32 x,y,z,r per point, KITTI layout)
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "vloam_adapter.hpp"

static vloam::CloudXYZ read_bin(const char* path) {
  vloam::CloudXYZ c;
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  float v[4];
  while (fread(v, sizeof(float), 4, f) == 4) { vloam::PointXYZ p; p.x = v[0]; p.y = v[1]; p.z = v[2]; c.points.push_back(p); }
  fclose(f);
  return c;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  vloam_b200_params prm;
  vloam_b200_default_params(&prm);
  prm.n_scans = 16; prm.minimum_range = 0.3f; prm.line_res = 0.2f; prm.plane_res = 0.4f;  // loam_velodyne_VLP_16.launch
  try {
    vloam::LidarOdometryMapping loam(prm);
    loam.init();
    for (int k = 1; k < argc; ++k) {
      const vloam::CloudXYZ cloud = read_bin(argv[k]);
      loam.reset();
      loam.scanRegistrationIO(cloud);
      loam.laserOdometryIO();
      loam.laserMappingIO();
      vloam::CloudPtr full, sharp, less, flat, lessflat;
      loam.scan_registration.output(full, sharp, less, flat, lessflat);
      vloam::Quat q; vloam::Vec3 t; vloam::CloudPtr cl, sl, fr; bool skip;
      loam.laser_odometry.output(q, t, cl, sl, fr, skip);
      vloam::Quat qm; vloam::Vec3 tm;
      loam.laser_mapping.output(qm, tm);
      // the per-point helpers of the reference's public interface against the device path
      vloam::CloudPtr reg;
      loam.laser_mapping.registeredFullCloud(reg);                       // LM.cpp:901-905 on the device
      if (reg->size() != full->size()) { printf("ERROR: registered cloud size\n"); return 1; }
      for (size_t i = 0; i < full->size(); i += 97) {
        vloam::PointType m, back, st, en;
        loam.laser_mapping.pointAssociateToMap(&full->points[i], &m);     // LM.cpp:154-164 on the host
        if (memcmp(&m, &reg->points[i], sizeof m) != 0) { printf("ERROR: pointAssociateToMap differs from the device at %zu\n", i); return 1; }
        loam.laser_mapping.pointAssociateTobeMapped(&m, &back);
        loam.laser_odometry.TransformToStart(&full->points[i], &st);
        loam.laser_odometry.TransformToEnd(&full->points[i], &en);
        const float tol = 1e-4f * (1.0f + fabsf(full->points[i].x) + fabsf(full->points[i].y) + fabsf(full->points[i].z));
        if (fabsf(back.x - full->points[i].x) > tol || fabsf(back.y - full->points[i].y) > tol || fabsf(back.z - full->points[i].z) > tol ||
            fabsf(en.x - full->points[i].x) > tol || fabsf(en.y - full->points[i].y) > tol || fabsf(en.z - full->points[i].z) > tol ||
            en.intensity != (float)(int)full->points[i].intensity || st.intensity != full->points[i].intensity) {
          printf("ERROR: per-point helper round trip at %zu\n", i); return 1;
        }
      }
      vloam::CloudXYZI near_far, kept;
      near_far.points = full->points;
      vloam::ScanRegistration::removeClosedPointCloud(near_far, kept, 10.0f);   // SR.cpp:107-141
      for (const auto& p : kept.points) if (p.x * p.x + p.y * p.y + p.z * p.z < 100.0f) { printf("ERROR: removeClosedPointCloud\n"); return 1; }
      printf("frame %d kept %zu sharp %zu less %zu flat %zu lessflat %zu | odom %.9f %.9f %.9f | map %.9f %.9f %.9f\n", k - 1, full->size(),
             sharp->size(), less->size(), flat->size(), lessflat->size(), t.x, t.y, t.z, tm.x, tm.y, tm.z);
    }
    // a bad scan_line must be refused like SR.cpp:58-61
    prm.n_scans = 48;
    bool threw = false;
    try { vloam::LidarOdometryMapping bad(prm); } catch (const vloam::AdapterError&) { threw = true; }
    if (!threw) { printf("ERROR: scan_line 48 accepted\n"); return 1; }
  } catch (const vloam::AdapterError& e) {
    printf("ERROR: %s\n", e.what());
    return 1;
  }
  printf("adapter ok\n");
  return 0;
}
