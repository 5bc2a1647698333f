// stand-in for <pcl/point_cloud.h>: points / width / height / is_dense / Ptr, as the reference uses them
#pragma once
#include <stdint.h>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud {
public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;  // (boost::shared_ptr before PCL 1.11: the adapter only uses Ptr(new ...))
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
};
}  // namespace pcl
