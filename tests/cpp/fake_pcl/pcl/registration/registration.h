// TEST INFRASTRUCTURE: the part of pcl::Registration<PointSource, PointTarget> (PCL 1.12, registration/registration.h) that the
// nodes and the adapter rely on: virtual setInputSource / setInputTarget, align() -> virtual computeTransformation(), the
// protected members the engines set. Restated from the published interface, not copied.
#pragma once
#include <Eigen/Core>
#include <pcl/point_cloud.h>

#include <string>
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  using Matrix4 = Eigen::Matrix4f;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  virtual ~Registration() = default;
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (input_) output = *input_;
    converged_ = false;
    final_transformation_ = Matrix4::Identity();
    computeTransformation(output, guess);
  }

 protected:
  std::string reg_name_;
  int max_iterations_ = 10;
  Matrix4 final_transformation_ = Matrix4::Identity();
  double transformation_epsilon_ = 0.0;
  double corr_dist_threshold_ = 1e30;
  bool converged_ = false;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
};
}  // namespace pcl
