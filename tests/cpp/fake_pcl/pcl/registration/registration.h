// TEST INFRASTRUCTURE: the part of pcl::Registration<PointSource, PointTarget> (PCL 1.12, registration/registration.h +
// impl/registration.hpp) that the nodes and the adapter rely on, INCLUDING the lazy host search tree: setInputTarget arms
// target_cloud_updated_, and the next align() -> initCompute() builds a kd-tree over the whole target on the host (for a
// 1 M-point map: hundreds of milliseconds in front of a 0.2 ms GPU solve — SURVEY.md section 8a row 1). The stub counts those
// builds (host_tree_builds()) so that a test can fail when the adapter triggers one. getFitnessScore is NON-virtual like
// PCL's and walks that tree. Restated from the published interface, not copied.
#pragma once
#include <Eigen/Core>
#include <pcl/point_cloud.h>

#include <limits>
#include <string>
namespace pcl {
namespace search {
template <typename PointT>
class KdTree {  // pcl::search::KdTree: setInputCloud builds the FLANN index
 public:
  using Ptr = std::shared_ptr<KdTree<PointT>>;
  void setInputCloud(const typename pcl::PointCloud<PointT>::ConstPtr& cloud) {
    cloud_ = cloud;
    builds_++;
  }
  int builds() const { return builds_; }
  bool hasCloud() const { return (bool)cloud_; }

 private:
  typename pcl::PointCloud<PointT>::ConstPtr cloud_;
  int builds_ = 0;
};
}  // namespace search

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
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  Registration() : tree_(new KdTree) {}
  virtual ~Registration() = default;
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) {
    source_cloud_updated_ = true;
    input_ = cloud;
  }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    if (cloud->points.empty()) return;  // PCL_ERROR + ignored
    target_ = cloud;
    target_cloud_updated_ = true;
  }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  // NON-virtual in PCL: through a base pointer THIS runs, on the host tree (graph_based_slam_component.cpp:231)
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    (void)max_range;
    host_fitness_calls_++;
    return tree_->hasCloud() ? 0.0 : std::numeric_limits<double>::max();
  }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!initCompute()) return;
    if (input_) output = *input_;
    converged_ = false;
    final_transformation_ = Matrix4::Identity();
    computeTransformation(output, guess);
  }
  int host_tree_builds() const { return tree_->builds(); }
  int host_fitness_calls() const { return host_fitness_calls_; }

 protected:
  bool initCompute() {
    if (!target_) return false;
    if (target_cloud_updated_ && !force_no_recompute_) {  // the lazy host kd-tree over the whole target
      tree_->setInputCloud(target_);
      target_cloud_updated_ = false;
    }
    return (bool)input_;
  }
  std::string reg_name_;
  KdTreePtr tree_;
  int max_iterations_ = 10;
  Matrix4 final_transformation_ = Matrix4::Identity();
  double transformation_epsilon_ = 0.0;
  double corr_dist_threshold_ = 1e30;
  bool converged_ = false;
  bool target_cloud_updated_ = true, source_cloud_updated_ = true, force_no_recompute_ = false;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int host_fitness_calls_ = 0;
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
};
}  // namespace pcl
