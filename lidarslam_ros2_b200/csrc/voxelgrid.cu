// K4 — VoxelGrid centroid downsample as a hash-and-reduce over the rank index.
// Replaces pcl::VoxelGrid<PointXYZI>::filter (PCL 1.12, external) at its reference call sites
// scanmatcher_component.cpp:266-269, 311-314, 325-328, 444-447; graph_based_slam_component.cpp:225-226;
// apps/align.cpp:66-75. Same leaf indexing as VoxelGridCovariance (voxel_grid_covariance_omp_impl.hpp:67-103,
// 218-223); one output point per occupied leaf = mean of x, y, z, intensity (downsample_all_data_ = true),
// emitted in ascending leaf index because the rank of a leaf in the occupancy bitmap IS its output slot.
// Algorithmic HBM bytes: N*16 (read) + M*16 (write).
#include "engine.hpp"

namespace b200 {

namespace {

__device__ __forceinline__ unsigned rank_of(const RankWord* __restrict__ table, int cell) {
  RankWord w = table[cell >> 5];
  return w.prefix + __popc(w.bits & ((1u << (cell & 31)) - 1u));
}

__global__ void __launch_bounds__(256) vg_mark_kernel(const float4* __restrict__ pts, size_t n, GridGeom g, RankWord* table,
                                                      int* cell_of_point) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int cell = -1;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    cell = build_leaf_index(g, p.x, p.y, p.z);
    if (cell < 0 || cell >= g.n_cells) cell = -1;
  }
  cell_of_point[i] = cell;
  if (cell >= 0) atomicOr(&table[cell >> 5].bits, 1u << (cell & 31));
}

__global__ void __launch_bounds__(256) vg_accumulate_kernel(const float4* __restrict__ pts, size_t n,
                                                            const int* __restrict__ cell_of_point,
                                                            const RankWord* __restrict__ table, double* acc) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cell = cell_of_point[i];
  if (cell < 0) return;
  double* a = acc + (size_t)rank_of(table, cell) * 5;
  float4 p = pts[i];
  atomicAdd(a + 0, (double)p.x);
  atomicAdd(a + 1, (double)p.y);
  atomicAdd(a + 2, (double)p.z);
  atomicAdd(a + 3, (double)p.w);
  atomicAdd(a + 4, 1.0);
}

__global__ void __launch_bounds__(256) vg_finalize_kernel(const double* __restrict__ acc, size_t m, float4* out) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= m) return;
  const double* a = acc + r * 5;
  const double inv = 1.0 / a[4];
  out[r] = make_float4((float)(a[0] * inv), (float)(a[1] * inv), (float)(a[2] * inv), (float)(a[3] * inv));
}

}  // namespace

long long VoxelGridFilter::filter_device(const float4* d_in, size_t n, float leaf, cudaStream_t s) {
  if (n == 0) return 0;
  bounds_scratch.ensure(8);
  Bounds b = cloud_bounds(d_in, n, bounds_scratch.ptr, s);
  launches += 1;
  if (!b.any) return 0;
  GridGeom g;
  if (!make_grid_geom(b, leaf, g)) return -1;  // PCL: "Leaf size is too small", output = input
  index.ensure((size_t)g.n_words);
  cell_of_point.ensure(n);
  rank_index_clear(index.ptr, g.n_words, s);
  const int blocks = (int)((n + 255) / 256);
  vg_mark_kernel<<<blocks, 256, 0, s>>>(d_in, n, g, index.ptr, cell_of_point.ptr);
  size_t m = rank_index_scan(index.ptr, g.n_words, scan_scratch, s);
  launches += 4;
  if (m == 0) return 0;
  acc.ensure(m * 5);
  out.ensure(m);
  B200_CUDA(cudaMemsetAsync(acc.ptr, 0, sizeof(double) * m * 5, s));
  vg_accumulate_kernel<<<blocks, 256, 0, s>>>(d_in, n, cell_of_point.ptr, index.ptr, acc.ptr);
  vg_finalize_kernel<<<(int)((m + 255) / 256), 256, 0, s>>>(acc.ptr, m, out.ptr);
  launches += 2;
  B200_CUDA(cudaGetLastError());
  return (long long)m;
}

}  // namespace b200
