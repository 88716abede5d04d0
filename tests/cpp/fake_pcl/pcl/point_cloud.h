// TEST INFRASTRUCTURE: pcl::PointCloud<PointT> as far as the adapter uses it (PCL 1.12 shapes).
#pragma once
#include <cstddef>
#include <memory>
#include <vector>
namespace pcl {
template <typename T>
using shared_ptr = std::shared_ptr<T>;
template <typename PointT>
class PointCloud {
 public:
  using Ptr = shared_ptr<PointCloud<PointT>>;
  using ConstPtr = shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
};
}  // namespace pcl
