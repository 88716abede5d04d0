// b200reg_pcl.hpp — C++ adapter over the C-ABI (b200reg.h) with the method names of pcl::Registration /
// pclomp::NormalDistributionsTransform / pclomp::GeneralizedIterativeClosestPoint, so that the two lidarslam_ros2 nodes
// change one `new` expression each (scanmatcher/src/scanmatcher_component.cpp:105-106,116-117;
// graph_based_slam/src/graph_based_slam_component.cpp:64-65,74-75). See INTEGRATION.md.
//
//  * With -DB200REG_WITH_PCL (a ROS 2 box with PCL >= 1.12) the classes derive from
//    pcl::Registration<PointSource, PointTarget>: the nodes keep holding them through
//    `boost::shared_ptr<pcl::Registration<pcl::PointXYZI, pcl::PointXYZI>> registration_`
//    (scanmatcher_component.h:93, graph_based_slam_component.h:106) and call the base-class API unchanged; the virtual
//    hook computeTransformation(output, guess) (ndt_omp.h:257-268, gicp_omp.h:332-333) forwards to b200reg_align().
//  * Without it (this repository's image has no PCL/Eigen) the same classes compile stand-alone over a minimal cloud type
//    with identical method names, which is what tests/cpp/adapter_smoke.cpp exercises.
// Header-only; link with -lb200reg. Matrices are column-major float[16] == Eigen::Matrix4f::data().
#pragma once
#include <array>
#include <cfloat>
#include <cstdio>
#include <limits>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200reg.h"

#ifdef B200REG_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>
#endif

namespace b200reg {

using Matrix4f = std::array<float, 16>;  // column-major, like Eigen::Matrix4f::data()
inline Matrix4f Identity() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }

// pclomp::NeighborSearchMethod (ndt_omp.h:52-57)
enum NeighborSearchMethod { KDTREE = B200REG_KDTREE, DIRECT26 = B200REG_DIRECT26, DIRECT7 = B200REG_DIRECT7, DIRECT1 = B200REG_DIRECT1 };

// RAII owner of a C-ABI handle
class Handle {
 public:
  Handle(int kind, int device) {
    if (b200reg_create(kind, device, &h_) != B200REG_OK)
      throw std::runtime_error("b200reg_create failed: no CUDA device (the engine has no CPU fallback)");
  }
  ~Handle() { b200reg_destroy(h_); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  b200reg_t get() const { return h_; }

 private:
  b200reg_t h_ = nullptr;
};

#ifndef B200REG_WITH_PCL
// Minimal stand-in for pcl::PointXYZI / pcl::PointCloud (same 32-byte layout: x y z 1 | intensity pad pad pad)
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, w = 1.0f;
  float intensity = 0, pad[3] = {0, 0, 0};
};
struct PointCloud {
  std::vector<PointXYZI> points;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
};

// pcl::Registration-shaped base: the surface the nodes exercise
class Registration {
 public:
  virtual ~Registration() = default;
  void setInputTarget(const PointCloud& cloud) {  // Registration::setInputTarget; empty clouds are ignored like PCL
    if (cloud.empty()) return;
    check(b200reg_set_input_target(h_.get(), &cloud.points[0].x, cloud.size(), sizeof(PointXYZI)));
  }
  void setInputSource(const PointCloud& cloud) {
    if (cloud.empty()) return;
    n_source_ = cloud.size();
    check(b200reg_set_input_source(h_.get(), &cloud.points[0].x, cloud.size(), sizeof(PointXYZI)));
  }
  void setTransformationEpsilon(double eps) { check(b200reg_set_transformation_epsilon(h_.get(), eps)); }
  void setMaximumIterations(int n) { check(b200reg_set_maximum_iterations(h_.get(), n)); }
  void setMaxCorrespondenceDistance(double d) { check(b200reg_set_max_correspondence_distance(h_.get(), d)); }
  void setEuclideanFitnessEpsilon(double e) { check(b200reg_set_euclidean_fitness_epsilon(h_.get(), e)); }
  void setRANSACIterations(int n) { check(b200reg_set_ransac_iterations(h_.get(), n)); }
  // align(output [, guess]): output receives the transformed source, like pcl::Registration::align
  void align(PointCloud& output, const Matrix4f& guess = Identity()) {
    int rc = b200reg_align(h_.get(), guess.data(), final_.data());
    if (rc != B200REG_OK && rc != B200REG_ERR_NO_TARGET && rc != B200REG_ERR_NO_SOURCE) check(rc);
    output.points.resize(n_source_);
    if (rc == B200REG_OK && n_source_) check(b200reg_get_aligned(h_.get(), &output.points[0].x, sizeof(PointXYZI)));
  }
  Matrix4f getFinalTransformation() const { return final_; }
  bool hasConverged() const {
    int c = 0;
    b200reg_has_converged(h_.get(), &c);
    return c != 0;
  }
  b200reg_t handle() const { return h_.get(); }  // for ScanMatcherSession
  double getFitnessScore(double max_range = DBL_MAX) {
    double v = DBL_MAX;
    check(b200reg_get_fitness_score(h_.get(), max_range, &v));
    return v;
  }

 protected:
  Registration(int kind, int device) : h_(kind, device) {}
  void check(int rc) const {
    if (rc != B200REG_OK) throw std::runtime_error(std::string("b200reg: ") + b200reg_last_error(h_.get()));
  }
  Handle h_;
  Matrix4f final_ = Identity();
  std::size_t n_source_ = 0;
};

class NormalDistributionsTransform : public Registration {
 public:
  explicit NormalDistributionsTransform(int device = 0) : Registration(B200REG_NDT, device) {}
  void setResolution(float r) { check(b200reg_ndt_set_resolution(h_.get(), r)); }
  void setStepSize(double s) { check(b200reg_ndt_set_step_size(h_.get(), s)); }
  void setOulierRatio(double r) { check(b200reg_ndt_set_outlier_ratio(h_.get(), r)); }  // sic, ndt_omp.h:180
  void setNeighborhoodSearchMethod(NeighborSearchMethod m) { check(b200reg_ndt_set_neighborhood_search_method(h_.get(), m)); }
  void setNumThreads(int n) { check(b200reg_ndt_set_num_threads(h_.get(), n)); }
  double getTransformationProbability() const {
    double v = 0;
    b200reg_ndt_get_transformation_probability(h_.get(), &v);
    return v;
  }
  int getFinalNumIteration() const {
    int v = 0;
    b200reg_ndt_get_final_num_iteration(h_.get(), &v);
    return v;
  }
};

class GeneralizedIterativeClosestPoint : public Registration {
 public:
  explicit GeneralizedIterativeClosestPoint(int device = 0) : Registration(B200REG_GICP, device) {}
  void setRotationEpsilon(double e) { check(b200reg_gicp_set_rotation_epsilon(h_.get(), e)); }
  void setCorrespondenceRandomness(int k) { check(b200reg_gicp_set_correspondence_randomness(h_.get(), k)); }
  void setMaximumOptimizerIterations(int n) { check(b200reg_gicp_set_maximum_optimizer_iterations(h_.get(), n)); }
};

#else  // ---------------------------------------------------------------------------------- B200REG_WITH_PCL

// Drop-in for pclomp::NormalDistributionsTransform<PointSource, PointTarget>: derives from pcl::Registration, so
// `registration_ = ndt;` (scanmatcher_component.cpp:113, graph_based_slam_component.cpp:72) keeps compiling.
template <typename PointSource, typename PointTarget>
class RegistrationBase : public pcl::Registration<PointSource, PointTarget> {
 protected:
  using Base = pcl::Registration<PointSource, PointTarget>;
  using typename Base::PointCloudSource;
  using typename Base::PointCloudTargetConstPtr;
  using typename Base::PointCloudSourceConstPtr;
  RegistrationBase(int kind, int device) : h_(kind, device) {}

 public:
  // pcl::Registration::setInputTarget arms target_cloud_updated_, and the next align() -> initCompute() then builds a
  // FLANN kd-tree over the WHOLE target on the host (hundreds of milliseconds for a 1 M-point map) that neither engine
  // here ever queries. So the target is stored WITHOUT arming that rebuild; the cloud goes to the GPU instead.
  // setKeepHostSearchTree(true) restores PCL's behaviour for callers that need PCL's own (non-virtual, host-side)
  // getFitnessScore through a base-class pointer.
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    if (!cloud || cloud->points.empty()) {
      PCL_ERROR_B200("[b200reg::setInputTarget] invalid or empty point cloud given, ignored");
      return;
    }
    if (keep_host_tree_) {
      Base::setInputTarget(cloud);
    } else {
      this->target_ = cloud;
      this->target_cloud_updated_ = false;  // (PCL constructs it `true`: even the first align() would build the tree)
    }
    target_ok_ = report(b200reg_set_input_target(h_.get(), &cloud->points[0].x, cloud->size(), sizeof(PointTarget)), "setInputTarget");
  }
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    Base::setInputSource(cloud);
    if (!cloud || cloud->points.empty()) {
      source_ok_ = false;
      return;
    }
    source_ok_ = report(b200reg_set_input_source(h_.get(), &cloud->points[0].x, cloud->size(), sizeof(PointSource)), "setInputSource");
  }
  void setKeepHostSearchTree(bool keep) { keep_host_tree_ = keep; }
  // align() fills `output` with the transformed source (a device-to-host copy of the whole scan per call). Both nodes
  // discard it (scanmatcher_component.cpp:350-358, graph_based_slam_component.cpp:229-231), so it is off by default:
  // `output` then keeps PCL's pre-filled copy of the input. Switch it on for callers that read the aligned cloud.
  void setComputeOutputCloud(bool on) { compute_output_ = on; }
  b200reg_t handle() const { return h_.get(); }  // for ScanMatcherSession
  // getFitnessScore on the GPU (exact 1-NN). pcl::Registration::getFitnessScore is NOT virtual: call this through the
  // concrete type, or through b200reg::getFitnessScore(registration_) below (gbs.cpp:231, sm.cpp:376).
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double v = std::numeric_limits<double>::max();
    report(b200reg_get_fitness_score(h_.get(), max_range, &v), "getFitnessScore");
    return v;
  }
  const char* lastError() const { return b200reg_last_error(h_.get()); }

 protected:
  // the virtual hook pcl::Registration::align() calls (ndt_omp.h:257-268, gicp_omp.h:332-333)
  void computeTransformation(PointCloudSource& output, const Eigen::Matrix4f& guess) override {
    this->converged_ = false;
    if (!target_ok_ || !source_ok_) {  // a failed upload must not be answered from the previous cloud
      PCL_ERROR_B200("[b200reg::align] the last setInputTarget / setInputSource failed; not aligning against stale data");
      return;
    }
    report(b200reg_set_transformation_epsilon(h_.get(), this->transformation_epsilon_), "setTransformationEpsilon");
    report(b200reg_set_maximum_iterations(h_.get(), this->max_iterations_), "setMaximumIterations");
    report(b200reg_set_max_correspondence_distance(h_.get(), this->corr_dist_threshold_), "setMaxCorrespondenceDistance");
    Eigen::Matrix4f final_t = Eigen::Matrix4f::Identity();
    const int rc = b200reg_align(h_.get(), guess.data(), final_t.data());
    this->final_transformation_ = final_t;
    int conv = 0;
    b200reg_has_converged(h_.get(), &conv);
    this->converged_ = report(rc, "align") && conv;
    if (rc == B200REG_OK && compute_output_ && !output.empty())
      report(b200reg_get_aligned(h_.get(), &output.points[0].x, sizeof(PointSource)), "getAligned");
  }
  bool report(int rc, const char* what) const {
    if (rc == B200REG_OK) return true;
    std::fprintf(stderr, "[b200reg::%s] error %d: %s\n", what, rc, b200reg_last_error(h_.get()));
    return false;
  }
  static void PCL_ERROR_B200(const char* msg) { std::fprintf(stderr, "%s\n", msg); }
  Handle h_;
  bool keep_host_tree_ = false, compute_output_ = false;
  bool target_ok_ = false, source_ok_ = false;
};

// getFitnessScore for code that only holds the base-class pointer (graph_based_slam_component.cpp:231,
// scanmatcher_component.cpp:376): GPU path for the engines of this header, PCL's own for anything else.
template <typename PointSource, typename PointTarget>
double getFitnessScore(pcl::Registration<PointSource, PointTarget>& reg, double max_range = std::numeric_limits<double>::max()) {
  if (auto* p = dynamic_cast<RegistrationBase<PointSource, PointTarget>*>(&reg)) return p->getFitnessScore(max_range);
  return reg.getFitnessScore(max_range);
}

template <typename PointSource, typename PointTarget>
class NormalDistributionsTransform : public RegistrationBase<PointSource, PointTarget> {
  using B = RegistrationBase<PointSource, PointTarget>;

 public:
  explicit NormalDistributionsTransform(int device = 0) : B(B200REG_NDT, device) {
    this->reg_name_ = "b200reg::NormalDistributionsTransform";
    this->transformation_epsilon_ = 0.1;  // ndt_omp_impl.hpp:71-72
    this->max_iterations_ = 35;
  }
  void setResolution(float r) { this->report(b200reg_ndt_set_resolution(this->h_.get(), r), "setResolution"); }
  void setStepSize(double s) { b200reg_ndt_set_step_size(this->h_.get(), s); }
  void setOulierRatio(double r) { b200reg_ndt_set_outlier_ratio(this->h_.get(), r); }
  void setNeighborhoodSearchMethod(NeighborSearchMethod m) { b200reg_ndt_set_neighborhood_search_method(this->h_.get(), m); }
  void setNumThreads(int n) { b200reg_ndt_set_num_threads(this->h_.get(), n); }
  double getTransformationProbability() const {
    double v = 0;
    b200reg_ndt_get_transformation_probability(this->h_.get(), &v);
    return v;
  }
  int getFinalNumIteration() const {
    int v = 0;
    b200reg_ndt_get_final_num_iteration(this->h_.get(), &v);
    return v;
  }
};

template <typename PointSource, typename PointTarget>
class GeneralizedIterativeClosestPoint : public RegistrationBase<PointSource, PointTarget> {
  using B = RegistrationBase<PointSource, PointTarget>;

 public:
  explicit GeneralizedIterativeClosestPoint(int device = 0) : B(B200REG_GICP, device) {
    this->reg_name_ = "b200reg::GeneralizedIterativeClosestPoint";
    this->max_iterations_ = 200;  // gicp_omp.h:117-119
    this->transformation_epsilon_ = 5e-4;
    this->corr_dist_threshold_ = 5.;
  }
  void setRotationEpsilon(double e) { b200reg_gicp_set_rotation_epsilon(this->h_.get(), e); }
  void setCorrespondenceRandomness(int k) { b200reg_gicp_set_correspondence_randomness(this->h_.get(), k); }
  void setMaximumOptimizerIterations(int n) { b200reg_gicp_set_maximum_optimizer_iterations(this->h_.get(), n); }
};

#endif  // B200REG_WITH_PCL

// ---- frontend session: device-resident map maintenance (b200sm_*, include/b200reg.h) -------------------------------
// What ScanMatcherComponent keeps per node instead of targeted_cloud_ / map_array_msg_.submaps[i].cloud on the host:
// one call per frame (receiveCloud) or the individual steps (setScan / updateMap). `Reg` is any of the registration
// classes above (it only needs the C handle).
class ScanMatcherSession {
 public:
  explicit ScanMatcherSession(int device = 0) {
    b200sm_t s = nullptr;
    if (b200sm_create(device, &s) != B200REG_OK) throw std::runtime_error("b200sm_create: no CUDA device (there is no CPU fallback)");
    s_.reset(s, [](b200sm_t p) { b200sm_destroy(p); });
  }
  void setParams(float vg_size_for_input, float vg_size_for_map, int num_targeted_cloud, double trans_for_mapupdate,
                 bool use_min_max_filter = false, double scan_min_range = 0.1, double scan_max_range = 100.0) {
    check(b200sm_set_params(s_.get(), vg_size_for_input, vg_size_for_map, num_targeted_cloud, trans_for_mapupdate,
                            use_min_max_filter ? 1 : 0, scan_min_range, scan_max_range));
  }
  void setInitialPose(const double position[3], const double quat_xyzw[4]) { check(b200sm_set_initial_pose(s_.get(), position, quat_xyzw)); }
  // points: n structs of `stride` bytes with x, y, z floats first and the intensity float at `intensity_offset` (or -1)
  // pose7 = position + quaternion (x, y, z, w); final16 column-major like Eigen::Matrix4f::data()
  bool receiveCloud(b200reg_t reg, const float* points, size_t n, size_t stride, long intensity_offset, double pose7[7], float final16[16]) {
    int updated = 0;
    check(b200sm_receive_cloud(s_.get(), reg, points, n, stride, intensity_offset, pose7, final16, &updated));
    return updated != 0;
  }
  size_t setScan(b200reg_t reg, const float* points, size_t n, size_t stride, long intensity_offset) {
    size_t m = 0;
    check(b200sm_set_scan(s_.get(), reg, points, n, stride, intensity_offset, &m));
    return m;
  }
  void updateMap(b200reg_t reg, const float final16[16], const double position[3], const double quat_xyzw[4], bool adopt_now = true) {
    check(b200sm_update_map(s_.get(), reg, final16, position, quat_xyzw, adopt_now ? 1 : 0));
  }
  size_t numSubmaps() const {
    size_t n = 0;
    b200sm_num_submaps(s_.get(), &n);
    return n;
  }
  b200sm_t handle() const { return s_.get(); }

 private:
  void check(int rc) const {
    if (rc != B200REG_OK) throw std::runtime_error(std::string("b200sm: ") + b200sm_last_error(s_.get()));
  }
  std::shared_ptr<b200sm_session> s_;
};

}  // namespace b200reg
