// TEST INFRASTRUCTURE (see Eigen/Core next to this file): pcl::PointXYZI with PCL's 32-byte layout.
#pragma once
namespace pcl {
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, data_w = 1.0f;
  float intensity = 0, pad[3] = {0, 0, 0};
};
}  // namespace pcl
