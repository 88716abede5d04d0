// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED: the reference has no tests / golden vectors
// and cannot be compiled here (needs PCL, Eigen, Boost, FLANN); this file is a dependency-free CPU
// restatement used as the checker and as the timed CPU baseline. It is never on the product path.
//
// Restates, function by function (paths under Thirdparty/ndt_omp_ros2/include/pclomp/):
//   VoxelGridCovariance::applyFilter          voxel_grid_covariance_omp_impl.hpp:48-370
//   VoxelGridCovariance::getNeighborhoodAtPoint{,7,1}   ...impl.hpp:373-442
//   VoxelGridCovariance::radiusSearch         voxel_grid_covariance_omp.h:470-499
//   NormalDistributionsTransform::computeTransformation ndt_omp_impl.hpp:80-171
//   ::computeDerivatives                      ndt_omp_impl.hpp:179-284
//   ::computeAngleDerivatives                 ndt_omp_impl.hpp:287-393
//   ::computePointDerivatives (f32 / f64)     ndt_omp_impl.hpp:396-438 / 441-479
//   ::updateDerivatives                       ndt_omp_impl.hpp:482-535
//   ::computeHessian / updateHessian          ndt_omp_impl.hpp:538-594 / 597-629
//   ::updateIntervalMT / trialValueSelectionMT / computeStepLengthMT   ndt_omp_impl.hpp:632-916
//   ::calculateScore                          ndt_omp_impl.hpp:919-953
//   pcl::Registration::align / getFitnessScore (PCL 1.12, external)
// It keeps the reference's data structures on purpose (std::map leaves, 7 finds per point, dense
// float 4x4 / 4x6 / 24x6 products, double accumulators, omp schedule(guided, 8)) so that it is a fair
// timing baseline for the OpenMP path.
#pragma once
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "kdtree.hpp"
#include "linalg.hpp"
#include "voxelgrid.hpp"

namespace oracle {

enum NeighborSearchMethod { KDTREE = 0, DIRECT26 = 1, DIRECT7 = 2, DIRECT1 = 3 };

struct Leaf {
  int nr_points = 0;
  double mean[3] = {0, 0, 0};
  float centroid[3] = {0, 0, 0};
  double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // starts at IDENTITY (voxel_grid_covariance_omp.h:101)
  double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double evecs[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double evals[3] = {0, 0, 0};
};

class VoxelGridCovariance {
 public:
  int min_points_per_voxel = 6;          // voxel_grid_covariance_omp.h:204
  double min_covar_eigvalue_mult = 0.01;  // voxel_grid_covariance_omp.h:205
  float leaf_size = 0;
  GridGeom geom{};
  std::map<size_t, Leaf> leaves;
  std::vector<P3> voxel_centroids;
  std::vector<int> voxel_centroids_leaf_indices;
  KdTree kdtree;

  void filter(const std::vector<P3>& input, float leaf) {
    leaf_size = leaf;
    leaves.clear();
    voxel_centroids.clear();
    voxel_centroids_leaf_indices.clear();
    geom = grid_geometry(input, leaf);
    if (geom.overflow) return;  // impl.hpp:79-84: warn, empty output
    // first pass (impl.hpp:209-264, filter_field_name_ empty, downsample_all_data_ = false)
    for (const P3& p : input) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      int idx = leaf_index(geom, p);
      Leaf& leaf_ref = leaves[(size_t)idx];
      double pt[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) leaf_ref.mean[a] += pt[a];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) leaf_ref.cov[a * 3 + b] += pt[a] * pt[b];
      leaf_ref.centroid[0] += p.x;
      leaf_ref.centroid[1] += p.y;
      leaf_ref.centroid[2] += p.z;
      ++leaf_ref.nr_points;
    }
    // second pass (impl.hpp:282-367), ascending leaf index
    for (auto& kv : leaves) {
      Leaf& l = kv.second;
      for (int a = 0; a < 3; a++) l.centroid[a] /= static_cast<float>(l.nr_points);
      double pt_sum[3] = {l.mean[0], l.mean[1], l.mean[2]};
      for (int a = 0; a < 3; a++) l.mean[a] /= l.nr_points;
      if (l.nr_points < min_points_per_voxel) continue;
      voxel_centroids.push_back({l.centroid[0], l.centroid[1], l.centroid[2]});
      voxel_centroids_leaf_indices.push_back((int)kv.first);
      const double n = l.nr_points;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
          l.cov[a * 3 + b] = (l.cov[a * 3 + b] - 2 * (pt_sum[a] * l.mean[b])) / n + l.mean[a] * l.mean[b];
      for (int a = 0; a < 9; a++) l.cov[a] *= (n - 1.0) / n;
      sym_eigen3(l.cov, l.evals, l.evecs);
      double ev[3] = {l.evals[0], l.evals[1], l.evals[2]};
      if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
        l.nr_points = -1;
        continue;
      }
      double min_ev = min_covar_eigvalue_mult * ev[2];
      if (ev[0] < min_ev) {
        ev[0] = min_ev;
        if (ev[1] < min_ev) ev[1] = min_ev;
        double D[9] = {ev[0], 0, 0, 0, ev[1], 0, 0, 0, ev[2]};
        double Vinv[9], tmp[9];
        mat3_inverse(l.evecs, Vinv);
        mat3_mul(l.evecs, D, tmp);
        mat3_mul(tmp, Vinv, l.cov);
      }
      for (int a = 0; a < 3; a++) l.evals[a] = ev[a];
      mat3_inverse(l.cov, l.icov);
      double mx = l.icov[0], mn = l.icov[0];
      for (int a = 1; a < 9; a++) {
        mx = std::max(mx, l.icov[a]);
        mn = std::min(mn, l.icov[a]);
      }
      if (mx == std::numeric_limits<float>::infinity() || mn == -std::numeric_limits<float>::infinity())
        l.nr_points = -1;
    }
    if (!voxel_centroids.empty()) kdtree.build(voxel_centroids);
  }

  // impl.hpp:373-404; rel = 3 x n offsets
  int neighborhood(const int* rel, int nrel, const P3& pt, std::vector<const Leaf*>& out) const {
    out.clear();
    int ijk[3] = {static_cast<int>(std::floor(pt.x / leaf_size)), static_cast<int>(std::floor(pt.y / leaf_size)),
                  static_cast<int>(std::floor(pt.z / leaf_size))};
    for (int ni = 0; ni < nrel; ni++) {
      const int* d = rel + 3 * ni;
      bool inside = true;
      for (int a = 0; a < 3; a++)
        if (!(geom.min_b[a] - ijk[a] <= d[a] && geom.max_b[a] - ijk[a] >= d[a])) inside = false;
      if (!inside) continue;
      int key = 0;
      for (int a = 0; a < 3; a++) key += (ijk[a] + d[a] - geom.min_b[a]) * geom.divb_mul[a];
      auto it = leaves.find((size_t)key);
      if (it != leaves.end() && it->second.nr_points >= min_points_per_voxel) out.push_back(&it->second);
    }
    return (int)out.size();
  }
  int neighborhood7(const P3& pt, std::vector<const Leaf*>& out) const {
    static const int rel[21] = {0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1};
    return neighborhood(rel, 7, pt, out);
  }
  int neighborhood1(const P3& pt, std::vector<const Leaf*>& out) const {
    static const int rel[3] = {0, 0, 0};
    return neighborhood(rel, 1, pt, out);
  }
  // pcl::getAllNeighborCellIndices(): the 26 non-centre cells.
  int neighborhood26(const P3& pt, std::vector<const Leaf*>& out) const {
    // C++11 guarantees this initialisation is thread-safe (the search runs inside an OpenMP parallel region)
    static const std::vector<int> rel = [] {
      std::vector<int> v;
      for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
          for (int dx = -1; dx <= 1; dx++) {
            if (dx == 0 && dy == 0 && dz == 0) continue;
            v.push_back(dx);
            v.push_back(dy);
            v.push_back(dz);
          }
      return v;
    }();
    return neighborhood(rel.data(), 26, pt, out);
  }
  // voxel_grid_covariance_omp.h:470-499 (FLANN radius result set keeps dist < r^2)
  int radius_search(const P3& pt, double radius, std::vector<const Leaf*>& out) const {
    out.clear();
    if (voxel_centroids.empty()) return 0;
    std::vector<int> idx;
    std::vector<float> d2;
    float r2 = static_cast<float>(radius * radius);
    kdtree.radius(pt, r2, idx, d2);
    for (size_t k = 0; k < idx.size(); k++) {
      if (!(d2[k] < r2)) continue;
      auto it = leaves.find((size_t)voxel_centroids_leaf_indices[idx[k]]);
      out.push_back(&it->second);
    }
    return (int)out.size();
  }
};

class NDT {
 public:
  // parameters (ndt_omp_impl.hpp:47-76)
  float resolution = 1.0f;
  double step_size = 0.1;
  double outlier_ratio = 0.55;
  double transformation_epsilon = 0.1;
  int max_iterations = 35;
  int search_method = DIRECT7;
  int num_threads = omp_get_max_threads();

  // results
  float final_transformation[16];  // row-major
  bool converged = false;
  int nr_iterations = 0;
  double trans_probability = 0;
  int n_evaluations = 0;  // number of computeDerivatives calls in the last align (instrumentation)

  std::vector<P3> target, input;
  VoxelGridCovariance target_cells;
  KdTree target_tree;  // pcl::Registration::tree_ (built lazily in align / getFitnessScore)
  bool target_tree_dirty = true;

  NDT() { set_identity(final_transformation); }

  void setInputTarget(const std::vector<P3>& cloud) {  // ndt_omp.h:117-122
    if (cloud.empty()) return;
    target = cloud;
    target_tree_dirty = true;
    init();
  }
  void setInputSource(const std::vector<P3>& cloud) { input = cloud; }
  void setResolution(float r) {  // ndt_omp.h:127-137
    if (resolution != r) {
      resolution = r;
      if (!input.empty() && !target.empty()) init();
    }
  }
  void init() { target_cells.filter(target, resolution); }  // ndt_omp.h:271-278

  // pcl::Registration::align(output, guess). guess row-major 4x4 float.
  void align(const float* guess, std::vector<P3>* output = nullptr) {
    converged = false;
    set_identity(final_transformation);
    if (target.empty() || input.empty()) return;
    if (target_tree_dirty) {  // initCompute(): exact kd-tree over ALL target points
      target_tree.build(target);
      target_tree_dirty = false;
    }
    std::vector<P3> out = input;
    computeTransformation(out, guess);
    if (output) *output = out;
  }

  // pcl::Registration::getFitnessScore(max_range)
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    if (target_tree_dirty) {
      target_tree.build(target);
      target_tree_dirty = false;
    }
    std::vector<P3> tr;
    transform_cloud(input, tr, final_transformation);
    double sum = 0;
    int nr = 0;
    std::vector<int> idx;
    std::vector<float> d2;
    for (const P3& p : tr) {
      if (target_tree.knn(p, 1, idx, d2) < 1) continue;
      if (d2[0] <= max_range) {
        sum += d2[0];
        nr++;
      }
    }
    return nr > 0 ? sum / nr : std::numeric_limits<double>::max();
  }

  // ---- exposed for unit tests -------------------------------------------------------------
  double gauss_d1 = 0, gauss_d2 = 0, gauss_d3 = 0;
  float j_ang[8][4], h_ang[16][4];
  double j_ang_d[8][3], h_ang_d[15][3];

  void init_gauss() {  // ndt_omp_impl.hpp:88-93
    double c1 = 10 * (1 - outlier_ratio);
    double c2 = outlier_ratio / std::pow((double)resolution, 3);
    gauss_d3 = -std::log(c2);
    gauss_d1 = -std::log(c1 + c2) - gauss_d3;
    gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - gauss_d3) / gauss_d1);
  }

  static void transform_cloud(const std::vector<P3>& in, std::vector<P3>& out, const float* T) {
    out.resize(in.size());
    for (size_t i = 0; i < in.size(); i++) {
      const P3 p = in[i];
      out[i].x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
      out[i].y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
      out[i].z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
    }
  }

  void computeAngleDerivatives(const double* p) {  // ndt_omp_impl.hpp:287-393
    double cx, cy, cz, sx, sy, sz;
    if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
    if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
    if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
    const double J[8][3] = {
        {-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy},  // a
        {cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy},    // b
        {-sy * cz, sy * sz, cy},                                       // c
        {sx * cy * cz, -sx * cy * sz, sx * sy},                        // d
        {-cx * cy * cz, cx * cy * sz, -cx * sy},                       // e
        {-cy * sz, -cy * cz, 0},                                       // f
        {cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0},          // g
        {sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0}};          // h
    // f64 vectors (used by computeHessian) — d1 z-component is -sy (impl.hpp:359)
    const double Hd[15][3] = {
        {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy},   // a2
        {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy},  // a3
        {cx * cy * cz, -cx * cy * sz, cx * sy},                        // b2
        {sx * cy * cz, -sx * cy * sz, sx * sy},                        // b3
        {-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0},          // c2
        {cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0},          // c3
        {-cy * cz, cy * sz, -sy},                                      // d1 (f64: -sy)
        {-sx * sy * cz, sx * sy * sz, sx * cy},                        // d2
        {cx * sy * cz, -cx * sy * sz, -cx * cy},                       // d3
        {sy * sz, sy * cz, 0},                                         // e1
        {-sx * cy * sz, -sx * cy * cz, 0},                             // e2
        {cx * cy * sz, cx * cy * cz, 0},                               // e3
        {-cy * cz, cy * sz, 0},                                        // f1
        {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0},         // f2
        {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0}};        // f3
    for (int r = 0; r < 8; r++) {
      for (int c = 0; c < 3; c++) {
        j_ang_d[r][c] = J[r][c];
        j_ang[r][c] = static_cast<float>(J[r][c]);
      }
      j_ang[r][3] = 0.0f;
    }
    for (int r = 0; r < 15; r++) {
      for (int c = 0; c < 3; c++) {
        h_ang_d[r][c] = Hd[r][c];
        h_ang[r][c] = static_cast<float>(Hd[r][c]);
      }
      h_ang[r][3] = 0.0f;
    }
    h_ang[6][2] = static_cast<float>(sy);  // f32 table keeps +sy (impl.hpp:381) — live-path quirk
    for (int c = 0; c < 4; c++) h_ang[15][c] = 0.0f;
  }

  // f32 (impl.hpp:396-438). pg: 4x6, ph: 24x6 (row-major)
  void computePointDerivatives(const double* x, float* pg, float* ph, bool compute_hessian = true) const {
    float x4[4] = {(float)x[0], (float)x[1], (float)x[2], 0.0f};
    float xj[8];
    for (int r = 0; r < 8; r++) {
      float s = 0;
      for (int c = 0; c < 4; c++) s += j_ang[r][c] * x4[c];
      xj[r] = s;
    }
    pg[1 * 6 + 3] = xj[0];
    pg[2 * 6 + 3] = xj[1];
    pg[0 * 6 + 4] = xj[2];
    pg[1 * 6 + 4] = xj[3];
    pg[2 * 6 + 4] = xj[4];
    pg[0 * 6 + 5] = xj[5];
    pg[1 * 6 + 5] = xj[6];
    pg[2 * 6 + 5] = xj[7];
    if (compute_hessian) {
      float xh[16];
      for (int r = 0; r < 16; r++) {
        float s = 0;
        for (int c = 0; c < 4; c++) s += h_ang[r][c] * x4[c];
        xh[r] = s;
      }
      const float a[4] = {0, xh[0], xh[1], 0}, b[4] = {0, xh[2], xh[3], 0}, c[4] = {0, xh[4], xh[5], 0};
      const float d[4] = {xh[6], xh[7], xh[8], 0}, e[4] = {xh[9], xh[10], xh[11], 0}, f[4] = {xh[12], xh[13], xh[14], 0};
      auto put = [&](int row0, int col, const float* v) {
        for (int k = 0; k < 4; k++) ph[(row0 + k) * 6 + col] = v[k];
      };
      put(12, 3, a); put(16, 3, b); put(20, 3, c);
      put(12, 4, b); put(16, 4, d); put(20, 4, e);
      put(12, 5, c); put(16, 5, e); put(20, 5, f);
    }
  }

  // f64 (impl.hpp:441-479). pg: 3x6, ph: 18x6
  void computePointDerivativesD(const double* x, double* pg, double* ph) const {
    auto dot = [&](const double* v) { return x[0] * v[0] + x[1] * v[1] + x[2] * v[2]; };
    pg[1 * 6 + 3] = dot(j_ang_d[0]);
    pg[2 * 6 + 3] = dot(j_ang_d[1]);
    pg[0 * 6 + 4] = dot(j_ang_d[2]);
    pg[1 * 6 + 4] = dot(j_ang_d[3]);
    pg[2 * 6 + 4] = dot(j_ang_d[4]);
    pg[0 * 6 + 5] = dot(j_ang_d[5]);
    pg[1 * 6 + 5] = dot(j_ang_d[6]);
    pg[2 * 6 + 5] = dot(j_ang_d[7]);
    const double a[3] = {0, dot(h_ang_d[0]), dot(h_ang_d[1])}, b[3] = {0, dot(h_ang_d[2]), dot(h_ang_d[3])};
    const double c[3] = {0, dot(h_ang_d[4]), dot(h_ang_d[5])};
    const double d[3] = {dot(h_ang_d[6]), dot(h_ang_d[7]), dot(h_ang_d[8])};
    const double e[3] = {dot(h_ang_d[9]), dot(h_ang_d[10]), dot(h_ang_d[11])};
    const double f[3] = {dot(h_ang_d[12]), dot(h_ang_d[13]), dot(h_ang_d[14])};
    auto put = [&](int row0, int col, const double* v) {
      for (int k = 0; k < 3; k++) ph[(row0 + k) * 6 + col] = v[k];
    };
    put(9, 3, a); put(12, 3, b); put(15, 3, c);
    put(9, 4, b); put(12, 4, d); put(15, 4, e);
    put(9, 5, c); put(12, 5, e); put(15, 5, f);
  }

  // impl.hpp:482-535
  double updateDerivatives(double* grad, double* hess, const float* pg, const float* ph, const double* x_trans,
                           const double* c_inv, bool compute_hessian) const {
    float x4[4] = {(float)x_trans[0], (float)x_trans[1], (float)x_trans[2], 0.0f};
    float c4[16] = {0};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) c4[r * 4 + c] = (float)c_inv[r * 3 + c];
    float gd2 = (float)gauss_d2;
    float xc[4];
    for (int j = 0; j < 4; j++) {
      float s = 0;
      for (int k = 0; k < 4; k++) s += x4[k] * c4[k * 4 + j];
      xc[j] = s;
    }
    float q = 0;
    for (int k = 0; k < 4; k++) q += x4[k] * xc[k];
    float e = std::exp(-gd2 * q * 0.5f);
    float score_inc = (float)(-gauss_d1 * e);
    e = gd2 * e;
    if (e > 1 || e < 0 || e != e) return 0;
    e = (float)(e * gauss_d1);
    float cg[24];  // 4x6 = c4 * pg
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 6; c++) {
        float s = 0;
        for (int k = 0; k < 4; k++) s += c4[r * 4 + k] * pg[k * 6 + c];
        cg[r * 6 + c] = s;
      }
    float v[6];
    for (int c = 0; c < 6; c++) {
      float s = 0;
      for (int k = 0; k < 4; k++) s += x4[k] * cg[k * 6 + c];
      v[c] = s;
    }
    for (int c = 0; c < 6; c++) grad[c] += (double)(e * v[c]);
    if (compute_hessian) {
      float pgcg[36];  // pg^T * cg  (6x6)
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
          float s = 0;
          for (int k = 0; k < 4; k++) s += pg[k * 6 + r] * cg[k * 6 + c];
          pgcg[r * 6 + c] = s;
        }
      for (int i = 0; i < 6; i++) {
        float h6[6];
        for (int j = 0; j < 6; j++) {
          float s = 0;
          for (int k = 0; k < 4; k++) s += xc[k] * ph[(i * 4 + k) * 6 + j];
          h6[j] = s;
        }
        for (int j = 0; j < 6; j++)
          hess[i * 6 + j] += (double)(e * (-gd2 * v[i] * v[j] + h6[j] + pgcg[j * 6 + i]));
      }
    }
    return score_inc;
  }

  // impl.hpp:179-284. trans_cloud = already transformed source cloud.
  double computeDerivatives(double* score_gradient, double* hessian, const std::vector<P3>& trans_cloud,
                            const double* p, bool compute_hessian = true) {
    n_evaluations++;
    const int nt = std::max(1, num_threads);
    std::vector<double> scores(nt, 0.0), grads((size_t)nt * 6, 0.0), hessians((size_t)nt * 36, 0.0);
    computeAngleDerivatives(p);
    std::vector<std::vector<const Leaf*>> neighborhoods(nt);
    const int n = (int)input.size();
#pragma omp parallel for num_threads(nt) schedule(guided, 8)
    for (int idx = 0; idx < n; idx++) {
      int tn = omp_get_thread_num();
      float pg[24], ph[144];
      std::memset(pg, 0, sizeof(pg));
      pg[0] = pg[7] = pg[14] = 1.0f;
      std::memset(ph, 0, sizeof(ph));
      const P3 x_trans_pt = trans_cloud[idx];
      auto& nb = neighborhoods[tn];
      switch (search_method) {
        case KDTREE: target_cells.radius_search(x_trans_pt, resolution, nb); break;
        case DIRECT26: target_cells.neighborhood26(x_trans_pt, nb); break;
        default:
        case DIRECT7: target_cells.neighborhood7(x_trans_pt, nb); break;
        case DIRECT1: target_cells.neighborhood1(x_trans_pt, nb); break;
      }
      double score_pt = 0, grad_pt[6] = {0}, hess_pt[36] = {0};
      for (const Leaf* cell : nb) {
        const P3 x_pt = input[idx];
        double x[3] = {x_pt.x, x_pt.y, x_pt.z};
        double xt[3] = {x_trans_pt.x - cell->mean[0], x_trans_pt.y - cell->mean[1], x_trans_pt.z - cell->mean[2]};
        computePointDerivatives(x, pg, ph);
        score_pt += updateDerivatives(grad_pt, hess_pt, pg, ph, xt, cell->icov, compute_hessian);
      }
      scores[tn] += score_pt;
      for (int k = 0; k < 6; k++) grads[(size_t)tn * 6 + k] += grad_pt[k];
      for (int k = 0; k < 36; k++) hessians[(size_t)tn * 36 + k] += hess_pt[k];
    }
    double score = 0;
    for (int k = 0; k < 6; k++) score_gradient[k] = 0;
    for (int k = 0; k < 36; k++) hessian[k] = 0;
    for (int t = 0; t < nt; t++) {
      score += scores[t];
      for (int k = 0; k < 6; k++) score_gradient[k] += grads[(size_t)t * 6 + k];
      for (int k = 0; k < 36; k++) hessian[k] += hessians[(size_t)t * 36 + k];
    }
    return score;
  }

  // impl.hpp:597-629
  void updateHessian(double* hess, const double* pg, const double* ph, const double* xt, const double* ci) const {
    auto cdot = [&](const double* v, double* out) {  // out = c_inv * v
      for (int r = 0; r < 3; r++) out[r] = ci[r * 3] * v[0] + ci[r * 3 + 1] * v[1] + ci[r * 3 + 2] * v[2];
    };
    double cx[3];
    cdot(xt, cx);
    double e = gauss_d2 * std::exp(-gauss_d2 * (xt[0] * cx[0] + xt[1] * cx[1] + xt[2] * cx[2]) / 2);
    if (e > 1 || e < 0 || e != e) return;
    e *= gauss_d1;
    for (int i = 0; i < 6; i++) {
      double col_i[3] = {pg[0 * 6 + i], pg[1 * 6 + i], pg[2 * 6 + i]}, cov_dxd_pi[3];
      cdot(col_i, cov_dxd_pi);
      for (int j = 0; j < 6; j++) {
        double col_j[3] = {pg[0 * 6 + j], pg[1 * 6 + j], pg[2 * 6 + j]}, cj[3], ch[3];
        cdot(col_j, cj);
        double hb[3] = {ph[(3 * i) * 6 + j], ph[(3 * i + 1) * 6 + j], ph[(3 * i + 2) * 6 + j]};
        cdot(hb, ch);
        auto d3 = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
        hess[i * 6 + j] += e * (-gauss_d2 * d3(xt, cov_dxd_pi) * d3(xt, cj) + d3(xt, ch) + d3(col_j, cov_dxd_pi));
      }
    }
  }

  // impl.hpp:538-594 (serial; radius neighbourhood regardless of search_method)
  void computeHessian(double* hessian, const std::vector<P3>& trans_cloud) {
    double pg[18] = {0}, ph[108] = {0};
    pg[0] = pg[7] = pg[14] = 1.0;
    for (int k = 0; k < 36; k++) hessian[k] = 0;
    std::vector<const Leaf*> nb;
    for (size_t idx = 0; idx < input.size(); idx++) {
      const P3 xtp = trans_cloud[idx];
      target_cells.radius_search(xtp, resolution, nb);
      for (const Leaf* cell : nb) {
        double x[3] = {input[idx].x, input[idx].y, input[idx].z};
        double xt[3] = {xtp.x - cell->mean[0], xtp.y - cell->mean[1], xtp.z - cell->mean[2]};
        computePointDerivativesD(x, pg, ph);
        updateHessian(hessian, pg, ph, xt, cell->icov);
      }
    }
  }

  // impl.hpp:632-670
  static bool updateIntervalMT(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u,
                               double a_t, double f_t, double g_t) {
    if (f_t > f_l) {
      a_u = a_t; f_u = f_t; g_u = g_t;
      return false;
    } else if (g_t * (a_l - a_t) > 0) {
      a_l = a_t; f_l = f_t; g_l = g_t;
      return false;
    } else if (g_t * (a_l - a_t) < 0) {
      a_u = a_l; f_u = f_l; g_u = g_l;
      a_l = a_t; f_l = f_t; g_l = g_t;
      return false;
    }
    return true;
  }

  // impl.hpp:673-753
  static double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u,
                                      double a_t, double f_t, double g_t) {
    if (f_t > f_l) {
      double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      double w = std::sqrt(z * z - g_t * g_l);
      double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
      if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
      return 0.5 * (a_q + a_c);
    } else if (g_t * g_l < 0) {
      double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      double w = std::sqrt(z * z - g_t * g_l);
      double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
      if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
      return a_s;
    } else if (std::fabs(g_t) <= std::fabs(g_l)) {
      double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      double w = std::sqrt(z * z - g_t * g_l);
      double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
      double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
      if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
      return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
    }
    double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }

  static double psiMT(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
  static double dpsiMT(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

  // impl.hpp:756-916
  double computeStepLengthMT(const double* x, double* step_dir, double step_init, double step_max, double step_min,
                             double& score, double* score_gradient, double* hessian, std::vector<P3>& trans_cloud) {
    double phi_0 = -score;
    double d_phi_0 = 0;
    for (int k = 0; k < 6; k++) d_phi_0 += score_gradient[k] * step_dir[k];
    d_phi_0 = -d_phi_0;
    double x_t[6];
    if (d_phi_0 >= 0) {
      if (d_phi_0 == 0) return 0;
      d_phi_0 *= -1;
      for (int k = 0; k < 6; k++) step_dir[k] *= -1;
    }
    const int max_step_iterations = 10;
    int step_iterations = 0;
    const double mu = 1.e-4, nu = 0.9;
    double a_l = 0, a_u = 0;
    double f_l = psiMT(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsiMT(d_phi_0, d_phi_0, mu);
    double f_u = psiMT(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsiMT(d_phi_0, d_phi_0, mu);
    // NB: true whenever step_max > step_min (impl.hpp:803) => the MT loop below is skipped
    bool interval_converged = (step_max - step_min) > 0, open_interval = true;
    double a_t = step_init;
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    for (int k = 0; k < 6; k++) x_t[k] = x[k] + step_dir[k] * a_t;
    pose_to_matrix_f(x_t, final_transformation);
    transform_cloud(input, trans_cloud, final_transformation);
    score = computeDerivatives(score_gradient, hessian, trans_cloud, x_t, true);
    auto dirdot = [&]() {
      double s = 0;
      for (int k = 0; k < 6; k++) s += score_gradient[k] * step_dir[k];
      return s;
    };
    double phi_t = -score, d_phi_t = -dirdot();
    double psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
    while (!interval_converged && step_iterations < max_step_iterations &&
           !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
      if (open_interval) a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
      else a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
      a_t = std::min(a_t, step_max);
      a_t = std::max(a_t, step_min);
      for (int k = 0; k < 6; k++) x_t[k] = x[k] + step_dir[k] * a_t;
      pose_to_matrix_f(x_t, final_transformation);
      transform_cloud(input, trans_cloud, final_transformation);
      score = computeDerivatives(score_gradient, hessian, trans_cloud, x_t, false);
      phi_t = -score;
      d_phi_t = -dirdot();
      psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu);
      d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
      if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
        open_interval = false;
        f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
        g_l = g_l + mu * d_phi_0;
        f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
        g_u = g_u + mu * d_phi_0;
      }
      if (open_interval) interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
      else interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
      step_iterations++;
    }
    if (step_iterations) computeHessian(hessian, trans_cloud);
    return a_t;
  }

  // impl.hpp:80-171. guess row-major 4x4 (nullptr = identity)
  void computeTransformation(std::vector<P3>& output, const float* guess) {
    nr_iterations = 0;
    converged = false;
    n_evaluations = 0;
    init_gauss();
    float ident[16];
    set_identity(ident);
    if (guess && std::memcmp(guess, ident, sizeof(ident)) != 0) {
      bool same = true;
      for (int k = 0; k < 16; k++) same = same && (guess[k] == ident[k]);
      if (!same) {
        std::memcpy(final_transformation, guess, sizeof(ident));
        transform_cloud(output, output, guess);
      }
    }
    float R[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) R[r * 3 + c] = final_transformation[r * 4 + c];
    float ang[3];
    euler_angles_012(R, ang);
    double p[6] = {final_transformation[3], final_transformation[7], final_transformation[11], ang[0], ang[1], ang[2]};
    double delta_p[6], score_gradient[6], hessian[36];
    double score = computeDerivatives(score_gradient, hessian, output, p);
    while (!converged) {
      JacobiSVD<6> sv(hessian);
      double neg_g[6];
      for (int k = 0; k < 6; k++) neg_g[k] = -score_gradient[k];
      sv.solve(neg_g, delta_p);
      double n2 = 0;
      for (int k = 0; k < 6; k++) n2 += delta_p[k] * delta_p[k];
      double delta_p_norm = std::sqrt(n2);
      if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
        trans_probability = score / static_cast<double>(input.size());
        converged = delta_p_norm == delta_p_norm;
        return;
      }
      for (int k = 0; k < 6; k++) delta_p[k] /= delta_p_norm;
      delta_p_norm = computeStepLengthMT(p, delta_p, delta_p_norm, step_size, transformation_epsilon / 2, score,
                                         score_gradient, hessian, output);
      for (int k = 0; k < 6; k++) delta_p[k] *= delta_p_norm;
      for (int k = 0; k < 6; k++) p[k] = p[k] + delta_p[k];
      if (nr_iterations > max_iterations || (nr_iterations && (std::fabs(delta_p_norm) < transformation_epsilon)))
        converged = true;
      nr_iterations++;
    }
    trans_probability = score / static_cast<double>(input.size());
  }

  // impl.hpp:919-953 (serial, radius neighbourhood, f64)
  double calculateScore(const std::vector<P3>& trans_cloud) {
    init_gauss();
    double score = 0;
    std::vector<const Leaf*> nb;
    for (const P3& xtp : trans_cloud) {
      target_cells.radius_search(xtp, resolution, nb);
      for (const Leaf* cell : nb) {
        double xt[3] = {xtp.x - cell->mean[0], xtp.y - cell->mean[1], xtp.z - cell->mean[2]};
        const double* ci = cell->icov;
        double cx[3];
        for (int r = 0; r < 3; r++) cx[r] = ci[r * 3] * xt[0] + ci[r * 3 + 1] * xt[1] + ci[r * 3 + 2] * xt[2];
        double e = std::exp(-gauss_d2 * (xt[0] * cx[0] + xt[1] * cx[1] + xt[2] * cx[2]) / 2);
        double score_inc = -gauss_d1 * e - gauss_d3;
        score += score_inc / nb.size();
      }
    }
    return score / static_cast<double>(trans_cloud.size());
  }

  static void set_identity(float* T) {
    for (int k = 0; k < 16; k++) T[k] = (k % 5 == 0) ? 1.0f : 0.0f;
  }
};

}  // namespace oracle
