// K4 — VoxelGrid centroid downsample as a hash-and-reduce over the rank index.
// Replaces pcl::VoxelGrid<PointXYZI>::filter (PCL 1.12, external) at its reference call sites
// scanmatcher_component.cpp:266-269, 311-314, 325-328, 444-447; graph_based_slam_component.cpp:225-226;
// apps/align.cpp:66-75. Same leaf indexing as VoxelGridCovariance (voxel_grid_covariance_omp_impl.hpp:67-103,
// 218-223); one output point per occupied leaf = mean of x, y, z, intensity (downsample_all_data_ = true),
// emitted in ascending leaf index because the rank of a leaf in the occupancy bitmap IS its output slot.
// Algorithmic HBM bytes: N*16 (read) + M*16 (write).
#include <algorithm>

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
  if (cell >= 0 && !((__ldcg(&table[cell >> 5].bits) >> (cell & 31)) & 1u)) atomicOr(&table[cell >> 5].bits, 1u << (cell & 31));
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


// ---- sparse (two-level) rank index --------------------------------------------------------------------------------
// The dense occupancy bitmap costs 8 bytes per 32 cells of the BOUNDING BOX: fine for a scan at 0.2 m, but a
// vg_size_for_map = 0.1 m filter over a 200 x 200 x 40 m map is 1.6e9 cells = 400 MB to clear and scan per call.
// pcl::VoxelGrid is O(N). Above a size budget the index therefore becomes two-level: level 1 is a rank index over PAGES of
// 1024 consecutive leaf indices (one bit per page), level 2 holds 32 RankWords only for the pages that are occupied,
// stored in page-rank order — at most min(N, pages) pages, i.e. O(N) memory — and one exclusive scan over the level-2
// words gives every leaf its output slot. The rank stays monotone in the leaf index, so the output order (ascending leaf
// index, the reference's sort order) is unchanged.
constexpr int PAGE_SHIFT = 10;                      // 1024 cells per page
constexpr int PAGE_WORDS = (1 << PAGE_SHIFT) / 32;  // 32 RankWords per occupied page

__global__ void __launch_bounds__(256) vg2_mark_pages_kernel(const float4* __restrict__ pts, size_t n, GridGeom g, RankWord* l1,
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
  if (cell >= 0) {
    const int page = cell >> PAGE_SHIFT;
    if (!((__ldcg(&l1[page >> 5].bits) >> (page & 31)) & 1u)) atomicOr(&l1[page >> 5].bits, 1u << (page & 31));
  }
}

__global__ void __launch_bounds__(256) vg2_mark_cells_kernel(size_t n, const int* __restrict__ cell_of_point,
                                                             const RankWord* __restrict__ l1, RankWord* l2) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = cell_of_point[i];
  if (cell < 0) return;
  const unsigned rp = rank_of(l1, cell >> PAGE_SHIFT);
  const int in_page = cell & ((1 << PAGE_SHIFT) - 1);
  unsigned* word = &l2[(size_t)rp * PAGE_WORDS + (in_page >> 5)].bits;
  if (!((__ldcg(word) >> (in_page & 31)) & 1u)) atomicOr(word, 1u << (in_page & 31));
}

__device__ __forceinline__ unsigned rank_of2(const RankWord* __restrict__ l1, const RankWord* __restrict__ l2, int cell) {
  const unsigned rp = rank_of(l1, cell >> PAGE_SHIFT);
  const int in_page = cell & ((1 << PAGE_SHIFT) - 1);
  const RankWord w = l2[(size_t)rp * PAGE_WORDS + (in_page >> 5)];
  return w.prefix + __popc(w.bits & ((1u << (in_page & 31)) - 1u));
}

__global__ void __launch_bounds__(256) vg2_accumulate_kernel(const float4* __restrict__ pts, size_t n,
                                                             const int* __restrict__ cell_of_point, const RankWord* __restrict__ l1,
                                                             const RankWord* __restrict__ l2, double* acc) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cell = cell_of_point[i];
  if (cell < 0) return;
  double* a = acc + (size_t)rank_of2(l1, l2, cell) * 5;
  float4 p = pts[i];
  atomicAdd(a + 0, (double)p.x);
  atomicAdd(a + 1, (double)p.y);
  atomicAdd(a + 2, (double)p.z);
  atomicAdd(a + 3, (double)p.w);
  atomicAdd(a + 4, 1.0);
}

__global__ void __launch_bounds__(256) vg_finalize_counted_kernel(const double* __restrict__ acc, const unsigned* __restrict__ m_ptr,
                                                                  float4* out) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= (size_t)*m_ptr) return;  // the number of occupied leaves stays on the device until the end of the call
  const double* a = acc + r * 5;
  const double inv = 1.0 / a[4];
  out[r] = make_float4((float)(a[0] * inv), (float)(a[1] * inv), (float)(a[2] * inv), (float)(a[3] * inv));
}

}  // namespace

long long VoxelGridFilter::filter_device(const float4* d_in, size_t n, float leaf, cudaStream_t s, const Bounds* known_bounds) {
  if (n == 0) return 0;
  Bounds b;
  if (known_bounds) {
    b = *known_bounds;
  } else {
    bounds_scratch.ensure(8);
    b = cloud_bounds(d_in, n, bounds_scratch.ptr, s);
    launches += 1;
  }
  if (!b.any) return 0;
  GridGeom g;
  if (!make_grid_geom(b, leaf, g)) return -1;  // PCL: "Leaf size is too small", output = input
  // buffers by the upper bound min(points, cells) on the occupied leaves; the count comes back once, at the end
  const size_t occ_max = (size_t)std::min<long long>((long long)n, g.n_cells);
  cell_of_point.ensure(n);
  acc.ensure(occ_max * 5);
  out.ensure(occ_max);
  count_dev.ensure(2);
  count_host.ensure(2);
  B200_CUDA(cudaMemsetAsync(acc.ptr, 0, sizeof(double) * occ_max * 5, s));
  const int blocks = (int)((n + 255) / 256);
  last_sparse = (size_t)g.n_words > dense_word_budget;
  if (!last_sparse) {
    index.ensure((size_t)g.n_words);
    rank_index_clear(index.ptr, g.n_words, s);
    vg_mark_kernel<<<blocks, 256, 0, s>>>(d_in, n, g, index.ptr, cell_of_point.ptr);
    rank_index_scan_async(index.ptr, g.n_words, scan_scratch, count_dev.ptr, s);
    vg_accumulate_kernel<<<blocks, 256, 0, s>>>(d_in, n, cell_of_point.ptr, index.ptr, acc.ptr);
    launches += 6;
  } else {
    const long long n_pages = (g.n_cells + (1 << PAGE_SHIFT) - 1) >> PAGE_SHIFT;
    const int l1_words = (int)((n_pages + 31) / 32);
    const size_t pages_max = (size_t)std::min<long long>((long long)n, n_pages);
    const size_t l2_words = pages_max * PAGE_WORDS;
    index.ensure((size_t)l1_words);
    index_l2.ensure(l2_words);
    rank_index_clear(index.ptr, l1_words, s);
    B200_CUDA(cudaMemsetAsync(index_l2.ptr, 0, sizeof(RankWord) * l2_words, s));
    vg2_mark_pages_kernel<<<blocks, 256, 0, s>>>(d_in, n, g, index.ptr, cell_of_point.ptr);
    rank_index_scan_async(index.ptr, l1_words, scan_scratch, count_dev.ptr + 1, s);
    vg2_mark_cells_kernel<<<blocks, 256, 0, s>>>(n, cell_of_point.ptr, index.ptr, index_l2.ptr);
    rank_index_scan_async(index_l2.ptr, (int)l2_words, scan_scratch, count_dev.ptr, s);
    vg2_accumulate_kernel<<<blocks, 256, 0, s>>>(d_in, n, cell_of_point.ptr, index.ptr, index_l2.ptr, acc.ptr);
    launches += 11;
  }
  vg_finalize_counted_kernel<<<(int)((occ_max + 255) / 256), 256, 0, s>>>(acc.ptr, count_dev.ptr, out.ptr);
  launches += 1;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(count_host.ptr, count_dev.ptr, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  return (long long)count_host.ptr[0];
}

}  // namespace b200
