// stand-in for <pcl/point_types.h>: the two point types of the path (common.h:42, scan_registration.h:71)
#pragma once
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
}  // namespace pcl
