// GICP engine (K5 kNN covariances, K6 correspondences + Mahalanobis, K7 cost/gradient reduction) behind the C-ABI.
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
  double gradient_tol = 1e-2;
};

struct GicpOutcome {
  float final_T[16];
  int converged, iterations, evaluations;
};

class GicpSolver {
 public:
  void init(int device, cudaStream_t s);
  void invalidate_target() { target_cov_valid_ = false; }
  void invalidate_source() { source_cov_valid_ = false; }
  GicpOutcome align(const NnGrid& target_grid, const float4* target, size_t n_target, const float4* source,
                    size_t n_source, const GicpConfig& cfg, const float* guess_rowmajor16, cudaStream_t s);
  int launches = 0;

 private:
  int device_ = 0;
  bool target_cov_valid_ = false, source_cov_valid_ = false;
};

}  // namespace b200
