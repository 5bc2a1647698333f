// main_node_excerpt.cpp -- the reference's ONLY caller of the lidar path, verbatim, compiled against vloam_adapter.hpp.
//
// The statements marked [MAIN.cpp:NNN] are copied character for character from
// /root/reference/src/vloam_main/src/vloam_main_node.cpp (lines 118, 120, 122, 124 of init(); 144 and 186-190 of
// callback(); the VisualOdometry statements between them are outside the hot path and left out).  If these compile and
// run against the adapter's headers, the drop-in claim of INTEGRATION.md holds for the reference's caller as written.
//
// Everything else here is the harness: the ROS parameter server is filled with the shipped launch-file values, sweeps
// are read from raw float files, and the poses the reference would save through vloam_tf (MAIN.cpp:193-198) are printed.
// usage: main_node_excerpt <scan_line> <minimum_range> <line_res> <plane_res> <skip> <scan0.bin> [scan1.bin ...]
#include <stdio.h>
#include <stdlib.h>
#include <memory>

#include <lidar_odometry_mapping/lidar_odometry_mapping.h>   // [MAIN.cpp:10]

std::shared_ptr<vloam::LidarOdometryMapping> LOAM;            // [MAIN.cpp:62]
pcl::PointCloud<pcl::PointXYZ> point_cloud_pcl;               // [MAIN.cpp:63]
std::shared_ptr<vloam::VloamTF> vloam_tf;                     // [MAIN.cpp:64]
int count;

void init() {
  count = 0;

  vloam_tf = std::make_shared<vloam::VloamTF>();               // [MAIN.cpp:118]
  LOAM = std::make_shared<vloam::LidarOdometryMapping>();      // [MAIN.cpp:120]

  vloam_tf->init();                                            // [MAIN.cpp:122]
  LOAM->init(vloam_tf);                                        // [MAIN.cpp:124]
}

void callback() {
  LOAM->reset();                                               // [MAIN.cpp:144]

  LOAM->scanRegistrationIO(point_cloud_pcl);                   // [MAIN.cpp:186]
  // 激光里程计计算连续两帧激光数据的位姿
  LOAM->laserOdometryIO();                                     // [MAIN.cpp:188]
  // 当前帧与submap匹配获得机器人在连续时刻的位置变换关系
  LOAM->laserMappingIO();                                      // [MAIN.cpp:190]

  ++count;
}

static void read_bin(const char* path, pcl::PointCloud<pcl::PointXYZ>& c) {
  c.clear();
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  float v[4];
  while (fread(v, sizeof(float), 4, f) == 4) { pcl::PointXYZ p; p.x = v[0]; p.y = v[1]; p.z = v[2]; c.push_back(p); }
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  // what loam_velodyne_*.launch + vloam_main.launch put on the parameter server
  ros::param::set("scan_line", atoi(argv[1]));
  ros::param::set("minimum_range", atof(argv[2]));
  ros::param::set("mapping_line_resolution", atof(argv[3]));
  ros::param::set("mapping_plane_resolution", atof(argv[4]));
  ros::param::set("mapping_skip_frame", atoi(argv[5]));
  ros::param::set("map_pub_number", 20);
  ros::param::set("loam_verbose_level", 1);
  ros::param::set("detach_VO_LO", true);
  try {
    init();
    for (int k = 6; k < argc; ++k) {
      read_bin(argv[k], point_cloud_pcl);
      callback();
      const tf2::Transform &lo = vloam_tf->world_LOT_base_last, &mo = vloam_tf->world_MOT_base_last, &ff = vloam_tf->base_prev_LOT_base_curr;
      printf("frame %d LO %.12f %.12f %.12f %.12f %.12f %.12f %.12f MO %.12f %.12f %.12f %.12f %.12f %.12f %.12f F2F %.12f %.12f %.12f\n", count - 1,
             lo.getRotation().x(), lo.getRotation().y(), lo.getRotation().z(), lo.getRotation().w(), lo.getOrigin().x(), lo.getOrigin().y(), lo.getOrigin().z(),
             mo.getRotation().x(), mo.getRotation().y(), mo.getRotation().z(), mo.getRotation().w(), mo.getOrigin().x(), mo.getOrigin().y(), mo.getOrigin().z(),
             ff.getOrigin().x(), ff.getOrigin().y(), ff.getOrigin().z());
    }
    // a missing parameter must stop init() like ROS_BREAK() (SR.cpp:50-54)
    ros::param::stub_store().erase("scan_line");
    bool threw = false;
    try { init(); } catch (const vloam::AdapterError&) { threw = true; }
    if (!threw) { printf("ERROR: init() without scan_line succeeded\n"); return 1; }
    ros::param::set("scan_line", 48);  // SR.cpp:58-61: only 16 / 32 / 64
    threw = false;
    try { init(); } catch (const vloam::AdapterError&) { threw = true; }
    if (!threw) { printf("ERROR: scan_line 48 accepted\n"); return 1; }
  } catch (const vloam::AdapterError& e) {
    printf("ERROR: %s\n", e.what());
    return 1;
  }
  printf("main node excerpt ok\n");
  return 0;
}
