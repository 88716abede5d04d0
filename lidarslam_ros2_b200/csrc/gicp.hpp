// GICP engine behind the C-ABI: K5 kNN covariances, K6 correspondences + Mahalanobis matrices, K7 cost / gradient
// reductions, and the host-side BFGS driver (pclomp::GeneralizedIterativeClosestPoint, gicp_omp_impl.hpp).
#pragma once
#include "engine.hpp"

namespace b200 {

constexpr int GICP_MAX_K = 32;

struct GicpConfig {  // gicp_omp.h:108-128
  int k_correspondences = 20;
  double gicp_epsilon = 0.001;
  double rotation_eps = 2e-3;
  int max_inner_iterations = 20;
  int max_iterations = 200;
  double trans_eps = 5e-4;
  double corr_dist = 5.0;
  double gradient_tol = 1e-2;  // see oracle/gicp.hpp header note on testGradient
};

struct GicpOutcome {
  float final_T[16];
  int converged, iterations, evaluations;
};

struct GicpInnerWork;     // device work area of the persistent inner-loop kernel (gicp.cu)
struct GicpInnerResult {  // written by the kernel into pinned host memory
  double x[6];
  double f;
  int status, inner, evaluations, error;
};

class GicpSolver {
 public:
  void init(int device, cudaStream_t s);
  void invalidate_target() { target_cov_valid_ = false; }
  void invalidate_source() {
    source_cov_valid_ = false;
    source_grid_valid_ = false;
  }
  GicpOutcome align(const NnGrid& target_grid, const float4* target, size_t n_target, const float4* source,
                    size_t n_source, const GicpConfig& cfg, const float* guess_rowmajor16, cudaStream_t s);
  // read-back for parity tests (row-major 3x3 doubles per point); which: 0 source, 1 target
  size_t covariances(int which, std::vector<double>& out, cudaStream_t s);
  int last_correspondences() const { return last_m_; }
  int launches = 0;
  // true: estimateRigidTransformationBFGS runs as ONE persistent cooperative kernel per outer iteration (BFGS on the
  // device); false: host-side BFGS, one K7 launch + synchronisation per functor evaluation
  bool device_bfgs = true;
  // accounting of the last align(): device time of the persistent inner kernel(s), their launches, and the
  // (correspondence, evaluation) products they processed (algorithmic bytes = that times 72)
  float inner_ms = 0;
  int inner_launches = 0;
  double inner_pair_evaluations = 0;

 private:
  void fdf(const float* T_rowmajor16, bool want_grad, double* f, double* g_t3, double* R9);
  // returns the BFGS status; x is updated in place
  int inner_loop_device(double* x, const GicpConfig& cfg, int* inner_iterations);
  GicpInnerWork* d_inner_work_ = nullptr;
  GicpInnerResult* h_inner_result_ = nullptr;  // pinned
  unsigned inner_epoch_ = 0;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  int sm_count_ = 0;
  int device_ = 0;
  cudaStream_t stream_ = nullptr;
  bool target_cov_valid_ = false, source_cov_valid_ = false, source_grid_valid_ = false;
  int cov_k_ = 0;
  double cov_eps_ = 0;
  NnGrid source_grid_;
  DeviceBuffer<double> target_cov_, source_cov_;  // 6 doubles per point (xx xy xz yy yz zz)
  DeviceBuffer<float> maha_;                      // 9 floats per source point
  DeviceBuffer<int> corr_;                        // target index per source point, -1 = none
  DeviceBuffer<int> nn_idx_;
  DeviceBuffer<float> nn_d2_;
  DeviceBuffer<float4> moved_;                    // source transformed by the guess ("output" cloud)
  DeviceBuffer<double> partials_;                 // per-CTA partial sums (16 doubles each)
  DeviceBuffer<double> result_;                   // 16 doubles
  DeviceBuffer<unsigned> counter_;
  double* h_result_ = nullptr;                    // pinned
  size_t n_source_ = 0, n_target_ = 0;
  const float4* target_ = nullptr;
  int last_m_ = 0;
  int evaluations_ = 0;
};

// k-NN based point covariances of a cloud against its own grid (gicp_omp_impl.hpp:48-122)
void gicp_covariances(const NnGrid& grid, const float4* pts, size_t n, int k, double gicp_epsilon, double* d_cov6,
                      cudaStream_t s);

}  // namespace b200
