// adapter_smoke.cpp -- the four classes of vloam_adapter.hpp driven stage by stage the way lidar_odometry_mapping.cpp:77-176
// drives the reference's (input / solve / publish / output with the reference's argument lists), plus the per-point helpers
// the reference exposes publicly, on sweeps read from raw float files.
// usage: adapter_smoke <scan0.bin> <scan1.bin>   (float32 x,y,z,r per point, KITTI layout)
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <lidar_odometry_mapping/scan_registration.h>
#include <lidar_odometry_mapping/laser_odometry.h>
#include <lidar_odometry_mapping/laser_mapping.h>

typedef pcl::PointCloud<vloam::PointType>::Ptr CloudPtr;

static pcl::PointCloud<pcl::PointXYZ> read_bin(const char* path) {
  pcl::PointCloud<pcl::PointXYZ> c;
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  float v[4];
  while (fread(v, sizeof(float), 4, f) == 4) { pcl::PointXYZ p; p.x = v[0]; p.y = v[1]; p.z = v[2]; c.push_back(p); }
  fclose(f);
  return c;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  // loam_velodyne_VLP_16.launch through an injected parameter source (no parameter server involved)
  vloam::adapter_param_source().get = [](const std::string& k, double& v) -> bool {
    if (k == "scan_line") v = 16; else if (k == "minimum_range") v = 0.3; else if (k == "mapping_line_resolution") v = 0.2;
    else if (k == "mapping_plane_resolution") v = 0.4; else if (k == "mapping_skip_frame") v = 1; else if (k == "map_pub_number") v = 20;
    else if (k == "loam_verbose_level") v = 0; else if (k == "detach_VO_LO") v = 1; else return false;
    return true;
  };
  try {
    std::shared_ptr<vloam::VloamTF> tf = std::make_shared<vloam::VloamTF>();
    tf->init();
    vloam::ScanRegistration sr;
    vloam::LaserOdometry lo;
    vloam::LaserMapping lm;
    sr.init();                      // creates the context
    lo.attach(sr.engine()); lm.attach(sr.engine());
    lo.init(tf); lm.init(tf);
    for (int k = 1; k < argc; ++k) {
      const pcl::PointCloud<pcl::PointXYZ> cloud = read_bin(argv[k]);
      sr.reset(); lm.reset();
      sr.input(cloud);
      CloudPtr full, sharp, less, flat, lessflat;
      sr.output(full, sharp, less, flat, lessflat);
      lo.input(full, sharp, less, flat, lessflat);
      lo.solveLO();
      lo.publish();
      Eigen::Quaterniond q; Eigen::Vector3d t; CloudPtr cl, sl, fr; bool skip = true;
      lo.output(q, t, cl, sl, fr, skip);
      lm.input(cl, sl, fr, q, t, skip);
      if (!skip) lm.solveMapping();
      lm.publish();
      Eigen::Quaterniond qm; Eigen::Vector3d tm;
      lm.output(qm, tm);
      if (tf->world_MOT_base_last.getOrigin().x() != tm.x() || tf->world_LOT_base_last.getOrigin().y() != t.y()) { printf("ERROR: VloamTF fields not written\n"); return 1; }
      // an edited cloud must be refused by input() (the device holds the real one)
      {
        CloudPtr edited(new pcl::PointCloud<vloam::PointType>(*sharp));
        if (!edited->points.empty()) edited->points[0].x += 1.0f;
        bool threw = edited->points.empty();
        try { lo.input(full, edited, less, flat, lessflat); } catch (const vloam::AdapterError&) { threw = true; }
        if (!threw) { printf("ERROR: input() accepted an edited cloud\n"); return 1; }
      }
      // the per-point helpers of the reference's public interface against the device path
      CloudPtr reg;
      lm.registeredFullCloud(reg);                       // LM.cpp:901-905 on the device
      if (reg->size() != full->size()) { printf("ERROR: registered cloud size\n"); return 1; }
      for (size_t i = 0; i < full->size(); i += 97) {
        vloam::PointType m, back, st, en;
        lm.pointAssociateToMap(&full->points[i], &m);     // LM.cpp:154-164 on the host
        if (memcmp(&m, &reg->points[i], sizeof m) != 0) { printf("ERROR: pointAssociateToMap differs from the device at %zu\n", i); return 1; }
        lm.pointAssociateTobeMapped(&m, &back);
        lo.TransformToStart(&full->points[i], &st);
        lo.TransformToEnd(&full->points[i], &en);
        const float tol = 1e-4f * (1.0f + fabsf(full->points[i].x) + fabsf(full->points[i].y) + fabsf(full->points[i].z));
        if (fabsf(back.x - full->points[i].x) > tol || fabsf(back.y - full->points[i].y) > tol || fabsf(back.z - full->points[i].z) > tol ||
            fabsf(en.x - full->points[i].x) > tol || fabsf(en.y - full->points[i].y) > tol || fabsf(en.z - full->points[i].z) > tol ||
            en.intensity != (float)(int)full->points[i].intensity || st.intensity != full->points[i].intensity) {
          printf("ERROR: per-point helper round trip at %zu\n", i); return 1;
        }
      }
      pcl::PointCloud<vloam::PointType> near_far(*full), kept;
      sr.removeClosedPointCloud(near_far, kept, 10.0f);   // SR.cpp:107-141
      for (const auto& p : kept.points) if (p.x * p.x + p.y * p.y + p.z * p.z < 100.0f) { printf("ERROR: removeClosedPointCloud\n"); return 1; }
      printf("frame %d kept %zu sharp %zu less %zu flat %zu lessflat %zu | odom %.9f %.9f %.9f | map %.9f %.9f %.9f\n", k - 1, full->size(),
             sharp->size(), less->size(), flat->size(), lessflat->size(), t.x(), t.y(), t.z(), tm.x(), tm.y(), tm.z());
    }
    // a bad scan_line must be refused like SR.cpp:58-61
    vloam::adapter_param_source().get = [](const std::string& k, double& v) -> bool { v = k == "scan_line" ? 48 : 1; return true; };
    bool threw = false;
    try { vloam::ScanRegistration bad; bad.init(); } catch (const vloam::AdapterError&) { threw = true; }
    if (!threw) { printf("ERROR: scan_line 48 accepted\n"); return 1; }
  } catch (const vloam::AdapterError& e) {
    printf("ERROR: %s\n", e.what());
    return 1;
  }
  printf("adapter ok\n");
  return 0;
}
