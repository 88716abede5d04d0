// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.12.1, external, not vendored;
// pinned by find_package(PCL 1.12 REQUIRED), Thirdparty/ndt_omp_ros2/CMakeLists.txt:23).
// Call sites in the reference: scanmatcher_component.cpp:266-269, 311-314, 325-328, 444-447;
// graph_based_slam_component.cpp:61, 225-226; apps/align.cpp:66-75.
// Restated from the published PCL algorithm: centroid of every field
// (x, y, z, intensity) per occupied leaf, float accumulators, output in ascending leaf index.
// The indexing (min_b/div_b/divb_mul, floor(p*inv_leaf) - (float)min_b) is the same in-tree code
// path as voxel_grid_covariance_omp_impl.hpp:67-103, 218-223.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace oracle {

struct P4 {
  float x, y, z, i;  // i = intensity
};

struct GridGeom {
  float inv_leaf[3];
  int min_b[3], max_b[3], div_b[3], divb_mul[3];
  bool overflow;
};

// Shared by VoxelGrid and VoxelGridCovariance (voxel_grid_covariance_omp_impl.hpp:67-103).
template <typename PT>
inline GridGeom grid_geometry(const std::vector<PT>& cloud, float leaf) {
  GridGeom g{};
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(),
                 std::numeric_limits<float>::max()};
  float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(),
                 -std::numeric_limits<float>::max()};
  for (const auto& p : cloud) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x);
    mn[1] = std::min(mn[1], p.y);
    mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x);
    mx[1] = std::max(mx[1], p.y);
    mx[2] = std::max(mx[2], p.z);
  }
  for (int a = 0; a < 3; a++) g.inv_leaf[a] = 1.0f / leaf;
  int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * g.inv_leaf[0]) + 1;
  int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * g.inv_leaf[1]) + 1;
  int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * g.inv_leaf[2]) + 1;
  g.overflow = (dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max());
  for (int a = 0; a < 3; a++) {
    g.min_b[a] = static_cast<int>(std::floor(mn[a] * g.inv_leaf[a]));
    g.max_b[a] = static_cast<int>(std::floor(mx[a] * g.inv_leaf[a]));
    g.div_b[a] = g.max_b[a] - g.min_b[a] + 1;
  }
  g.divb_mul[0] = 1;
  g.divb_mul[1] = g.div_b[0];
  g.divb_mul[2] = g.div_b[0] * g.div_b[1];
  return g;
}

template <typename PT>
inline int leaf_index(const GridGeom& g, const PT& p) {
  int ijk0 = static_cast<int>(std::floor(p.x * g.inv_leaf[0]) - static_cast<float>(g.min_b[0]));
  int ijk1 = static_cast<int>(std::floor(p.y * g.inv_leaf[1]) - static_cast<float>(g.min_b[1]));
  int ijk2 = static_cast<int>(std::floor(p.z * g.inv_leaf[2]) - static_cast<float>(g.min_b[2]));
  return ijk0 * g.divb_mul[0] + ijk1 * g.divb_mul[1] + ijk2 * g.divb_mul[2];
}

// pcl::VoxelGrid::filter with downsample_all_data_ = true, min_points_per_voxel_ = 0.
inline void voxelgrid_filter(const std::vector<P4>& in, float leaf, std::vector<P4>& out) {
  out.clear();
  if (in.empty()) return;
  GridGeom g = grid_geometry(in, leaf);
  if (g.overflow) {  // PCL warns and returns the input unchanged
    out = in;
    return;
  }
  struct Ref {
    int idx;
    int pt;
  };
  std::vector<Ref> refs;
  refs.reserve(in.size());
  for (size_t k = 0; k < in.size(); k++) {
    const P4& p = in[k];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    refs.push_back({leaf_index(g, p), (int)k});
  }
  // PCL uses an unstable integer sort on idx; the within-leaf order is unspecified. Stable here.
  std::stable_sort(refs.begin(), refs.end(), [](const Ref& a, const Ref& b) { return a.idx < b.idx; });
  size_t k = 0;
  while (k < refs.size()) {
    size_t e = k;
    float sx = 0, sy = 0, sz = 0, si = 0;
    while (e < refs.size() && refs[e].idx == refs[k].idx) {
      const P4& p = in[refs[e].pt];
      sx += p.x;
      sy += p.y;
      sz += p.z;
      si += p.i;
      e++;
    }
    float n = static_cast<float>(e - k);
    out.push_back({sx / n, sy / n, sz / n, si / n});
    k = e;
  }
}

}  // namespace oracle
