// K8 — exact nearest neighbour over the target cloud on a uniform cell grid (rank index + cell lists), and the
// fitness reduction. Replaces pcl::KdTreeFLANN::nearestKSearch(k=1) as used by
// pcl::Registration::getFitnessScore (called at graph_based_slam_component.cpp:231, scanmatcher_component.cpp:376,
// apps/align.cpp:37) and by GICP's searchForNeighbors (gicp_omp.h:340-347).
//
// Exactness: cells are visited in Chebyshev rings around the query's cell; after ring r every unvisited point is at
// least r*h away, so the search stops as soon as best_d2 <= (r*h)^2 (with a conservative float margin). Squared
// distances are accumulated in f32 exactly like FLANN's L2_Simple, ((dx*dx + dy*dy) + dz*dz), un-fused; ties go to
// the lower point index.
#include <cfloat>
#include <cmath>

#include "engine.hpp"
#include "nn_search.cuh"

namespace b200 {

namespace {

__device__ __forceinline__ unsigned rank_of(const RankWord* __restrict__ table, int cell) {
  RankWord w = table[cell >> 5];
  return w.prefix + __popc(w.bits & ((1u << (cell & 31)) - 1u));
}

__global__ void __launch_bounds__(256) nn_mark_kernel(const float4* __restrict__ pts, size_t n, NnGeom g, RankWord* table,
                                                      int* cell_of_point) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int cell = -1;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    int ix = nn_cell_coord(p.x, g.origin[0], g.inv_h, g.dims[0]);
    int iy = nn_cell_coord(p.y, g.origin[1], g.inv_h, g.dims[1]);
    int iz = nn_cell_coord(p.z, g.origin[2], g.inv_h, g.dims[2]);
    cell = ix + g.dims[0] * (iy + g.dims[1] * iz);
    atomicOr(&table[cell >> 5].bits, 1u << (cell & 31));
  }
  cell_of_point[i] = cell;
}

__global__ void __launch_bounds__(256) nn_count_kernel(size_t n, const int* __restrict__ cell_of_point,
                                                       const RankWord* __restrict__ table, unsigned* counts) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cell = cell_of_point[i];
  if (cell < 0) return;
  atomicAdd(&counts[rank_of(table, cell)], 1u);
}

__global__ void __launch_bounds__(256) nn_scatter_kernel(const float4* __restrict__ pts, size_t n,
                                                         const int* __restrict__ cell_of_point,
                                                         const RankWord* __restrict__ table,
                                                         const unsigned* __restrict__ cell_start, unsigned* cursor,
                                                         float4* sorted) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cell = cell_of_point[i];
  if (cell < 0) return;
  unsigned r = rank_of(table, cell);
  unsigned pos = cell_start[r] + atomicAdd(&cursor[r], 1u);
  float4 p = pts[i];
  sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float((int)i));
}

// exclusive scan of a u32 array in place (data[n] receives the total): tile-local scan, scan of the tile sums by one
// block, then the tile offsets are added back
constexpr int USCAN_THREADS = 256, USCAN_ITEMS = 8, USCAN_TILE = USCAN_THREADS * USCAN_ITEMS;

__global__ void __launch_bounds__(USCAN_THREADS) uscan_local_kernel(unsigned* data, size_t n, unsigned* tile_sums) {
  __shared__ unsigned warp_tot[USCAN_THREADS / 32];
  const size_t base = (size_t)blockIdx.x * USCAN_TILE + (size_t)threadIdx.x * USCAN_ITEMS;
  unsigned v[USCAN_ITEMS], local = 0;
#pragma unroll
  for (int k = 0; k < USCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? data[base + k] : 0u;
    local += v[k];
  }
  unsigned incl = local;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned off = 0;
  for (int w = 0; w < warp; w++) off += warp_tot[w];
  unsigned run = off + incl - local;
#pragma unroll
  for (int k = 0; k < USCAN_ITEMS; k++) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == USCAN_THREADS - 1) tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(1024) uscan_tiles_kernel(unsigned* tile_sums, int n_tiles, unsigned* total_out) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_tiles; base += 1024) {
    int i = base + threadIdx.x;
    unsigned v = (i < n_tiles) ? tile_sums[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned off = 0;
    for (int w = 0; w < warp; w++) off += warp_tot[w];
    unsigned carry = carry_s;
    if (i < n_tiles) tile_sums[i] = carry + off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

__global__ void __launch_bounds__(USCAN_THREADS) uscan_apply_kernel(unsigned* data, size_t n, const unsigned* tile_sums) {
  const unsigned off = tile_sums[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * USCAN_TILE + (size_t)threadIdx.x * USCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < USCAN_ITEMS; k++)
    if (base + k < n) data[base + k] += off;
}

constexpr int NN_MAX_RINGS = 3;  // 7^3 cells; beyond that a query is an outlier and goes to the brute-force pass

struct NnQueryParams {
  NnView V;
  float T[12];
  int has_T;
  float max_d2;
  int n_points;
};

__device__ __forceinline__ void nn_query_point(const NnQueryParams& P, float4 q4, float& qx, float& qy, float& qz) {
  qx = q4.x; qy = q4.y; qz = q4.z;
  if (P.has_T) {  // same un-fused float transform as the solver / pcl::transformPointCloud
    const float* T = P.T;
    const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], qx), __fmul_rn(T[1], qy)), __fmul_rn(T[2], qz)), T[3]);
    const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], qx), __fmul_rn(T[5], qy)), __fmul_rn(T[6], qz)), T[7]);
    const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], qx), __fmul_rn(T[9], qy)), __fmul_rn(T[10], qz)), T[11]);
    qx = tx; qy = ty; qz = tz;
  }
}

__global__ void __launch_bounds__(128) nn1_kernel(NnQueryParams P, const float4* __restrict__ queries, size_t n, int* out_idx,
                                                  float* out_d2, unsigned* unresolved_count, int* unresolved_list) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float qx, qy, qz;
  nn_query_point(P, queries[i], qx, qy, qz);
  float best;
  int best_i;
  const bool resolved = nn1_search(P.V, qx, qy, qz, P.max_d2, NN_MAX_RINGS, best, best_i);
  out_idx[i] = best_i;
  out_d2[i] = best;
  if (!resolved) unresolved_list[atomicAdd(unresolved_count, 1u)] = (int)i;
}

// Phase 2: one CTA per unresolved query. With a finite search radius the CTA shares the cells of the cube of that
// radius; without one it scans the whole (cell-sorted) cloud. Lexicographic (d2, index) minimum, so the result does
// not depend on the visiting order.
__global__ void __launch_bounds__(256) nn1_bruteforce_kernel(NnQueryParams P, const float4* __restrict__ queries,
                                                             const unsigned* __restrict__ unresolved_count,
                                                             const int* __restrict__ unresolved_list, int* out_idx,
                                                             float* out_d2) {
  __shared__ float sd[8];
  __shared__ int si[8];
  const unsigned total = *unresolved_count;
  for (unsigned u = blockIdx.x; u < total; u += gridDim.x) {
    const int qi = unresolved_list[u];
    float qx, qy, qz;
    nn_query_point(P, queries[qi], qx, qy, qz);
    float best = FLT_MAX;
    int best_i = -1;
    auto consider = [&](float4 t) {
      const float d2 = nn_dist2(qx, qy, qz, t);
      const int ti = __float_as_int(t.w);
      if (d2 < best || (d2 == best && ti < best_i)) {
        best = d2;
        best_i = ti;
      }
    };
    if (P.max_d2 < 3.0e38f) {
      // the caller only wants neighbours within sqrt(max_d2): the CTA shares the cells of the cube of that radius
      // around the query's cell (every point within the radius lies in it, also for queries outside the grid)
      const NnGeom& g = P.V.g;
      const int R = (int)ceilf(sqrtf(P.max_d2) * g.inv_h) + 1;
      const int cx = nn_cell_coord(qx, g.origin[0], g.inv_h, g.dims[0]);
      const int cy = nn_cell_coord(qy, g.origin[1], g.inv_h, g.dims[1]);
      const int cz = nn_cell_coord(qz, g.origin[2], g.inv_h, g.dims[2]);
      const int x0 = max(cx - R, 0), x1 = min(cx + R, g.dims[0] - 1);
      const int y0 = max(cy - R, 0), y1 = min(cy + R, g.dims[1] - 1);
      const int z0 = max(cz - R, 0), z1 = min(cz + R, g.dims[2] - 1);
      const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
      const long long cells = (long long)nx * ny * nz;
      for (long long c = threadIdx.x; c < cells; c += blockDim.x) {
        const int x = x0 + (int)(c % nx), y = y0 + (int)((c / nx) % ny), z = z0 + (int)(c / ((long long)nx * ny));
        const int cell = x + g.dims[0] * (y + g.dims[1] * z);
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(P.V.index + (cell >> 5)));
        const unsigned bit = cell & 31;
        if (!((w.x >> bit) & 1u)) continue;
        const unsigned rk = w.y + __popc(w.x & ((1u << bit) - 1u));
        const unsigned st = __ldg(P.V.cell_start + rk), en = __ldg(P.V.cell_start + rk + 1);
        for (unsigned k = st; k < en; k++) consider(__ldg(P.V.sorted + k));
      }
    } else {
      for (int k = threadIdx.x; k < P.n_points; k += blockDim.x) consider(__ldg(P.V.sorted + k));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, d);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, d);
      if (oi >= 0 && (od < best || (od == best && (best_i < 0 || oi < best_i)))) {
        best = od;
        best_i = oi;
      }
    }
    if ((threadIdx.x & 31) == 0) {
      sd[threadIdx.x >> 5] = best;
      si[threadIdx.x >> 5] = best_i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++)
        if (si[w] >= 0 && (sd[w] < best || (sd[w] == best && (best_i < 0 || si[w] < best_i)))) {
          best = sd[w];
          best_i = si[w];
        }
      out_idx[qi] = best_i;
      out_d2[qi] = best;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) fitness_kernel(const float* __restrict__ d2, const int* __restrict__ idx, size_t n,
                                                      double max_range, double* out2) {
  double s = 0, c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (idx[i] < 0) continue;
    double d = (double)d2[i];
    if (d <= max_range) {
      s += d;
      c += 1.0;
    }
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, k);
    c += __shfl_xor_sync(0xffffffffu, c, k);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out2, s);
    atomicAdd(out2 + 1, c);
  }
}

}  // namespace

NnView nn_view(const NnGrid& grid) {
  NnView V;
  V.index = grid.index.ptr;
  V.cell_start = grid.cell_start.ptr;
  V.sorted = grid.sorted.ptr;
  for (int a = 0; a < 3; a++) {
    V.g.origin[a] = grid.origin[a];
    V.g.dims[a] = grid.dims[a];
  }
  V.g.h = grid.h;
  V.g.inv_h = grid.inv_h;
  return V;
}

void NnGrid::build(const float4* pts, size_t n, cudaStream_t s, const Bounds* known_bounds) {
  valid = false;
  n_points = n;
  n_cells_occupied = 0;
  if (n == 0) return;
  Bounds b;
  if (known_bounds) {
    b = *known_bounds;
  } else {
    bounds_scratch.ensure(8);
    b = cloud_bounds(pts, n, bounds_scratch.ptr, s);
    launches += 1;
  }
  if (!b.any) return;
  double ext[3];
  for (int a = 0; a < 3; a++) {
    origin[a] = b.mn[a];
    ext[a] = std::max(1e-3, (double)b.mx[a] - (double)b.mn[a]);
  }
  double vol = ext[0] * ext[1] * ext[2];
  double hh = std::cbrt(vol / (4.0 * (double)n));
  hh = std::max(hh, 1e-3);
  for (;;) {  // keep the cell count below 2^28
    double cells = 1;
    for (int a = 0; a < 3; a++) cells *= std::floor(ext[a] / hh) + 1;
    if (cells <= 268435456.0) break;
    hh *= 1.26;
  }
  h = (float)hh;
  inv_h = 1.0f / h;
  n_cells = 1;
  for (int a = 0; a < 3; a++) {
    dims[a] = (int)std::floor(ext[a] / hh) + 1;
    n_cells *= dims[a];
  }
  n_words = (int)((n_cells + 31) / 32);
  // No host round trip below: the cell lists are sized by the upper bound min(points, cells) on the occupied cells; the
  // exclusive scan runs over that many counters (the tail past the occupied cells is zero, so cell_start[rank + 1] of
  // the last occupied cell is the total, as the queries expect).
  const size_t occ_max = (size_t)std::min<long long>((long long)n, n_cells);
  index.ensure((size_t)n_words);
  cell_of_point.ensure(n);
  cell_start.ensure(occ_max + 1);
  cursor.ensure(occ_max);
  sorted.ensure(n);
  scan_scratch.total.ensure(1);
  rank_index_clear(index.ptr, n_words, s);
  B200_CUDA(cudaMemsetAsync(cell_start.ptr, 0, sizeof(unsigned) * (occ_max + 1), s));
  B200_CUDA(cudaMemsetAsync(cursor.ptr, 0, sizeof(unsigned) * occ_max, s));
  NnGeom g;
  for (int a = 0; a < 3; a++) {
    g.origin[a] = origin[a];
    g.dims[a] = dims[a];
  }
  g.h = h;
  g.inv_h = inv_h;
  const int blocks = (int)((n + 255) / 256);
  nn_mark_kernel<<<blocks, 256, 0, s>>>(pts, n, g, index.ptr, cell_of_point.ptr);
  rank_index_scan_async(index.ptr, n_words, scan_scratch, scan_scratch.total.ptr, s);
  nn_count_kernel<<<blocks, 256, 0, s>>>(n, cell_of_point.ptr, index.ptr, cell_start.ptr);
  {
    const int n_tiles = (int)((occ_max + USCAN_TILE - 1) / USCAN_TILE);
    scan_tmp.ensure((size_t)n_tiles + 1);
    uscan_local_kernel<<<n_tiles, USCAN_THREADS, 0, s>>>(cell_start.ptr, occ_max, scan_tmp.ptr);
    uscan_tiles_kernel<<<1, 1024, 0, s>>>(scan_tmp.ptr, n_tiles, cell_start.ptr + occ_max);
    if (n_tiles > 1) uscan_apply_kernel<<<n_tiles, USCAN_THREADS, 0, s>>>(cell_start.ptr, occ_max, scan_tmp.ptr);
  }
  nn_scatter_kernel<<<blocks, 256, 0, s>>>(pts, n, cell_of_point.ptr, index.ptr, cell_start.ptr, cursor.ptr, sorted.ptr);
  launches += 9;
  B200_CUDA(cudaGetLastError());
  n_cells_occupied = occ_max;  // upper bound; the exact count stays on the device (scan_scratch.total)
  valid = true;
}

void nn1_query(const NnGrid& grid, const float4* queries, size_t n, const float* T12_host, int* d_idx, float* d_d2,
               cudaStream_t s, float max_d2) {
  if (n == 0) return;
  NnQueryParams P;
  P.V = nn_view(grid);
  P.has_T = T12_host ? 1 : 0;
  for (int k = 0; k < 12; k++) P.T[k] = T12_host ? T12_host[k] : 0.f;
  P.max_d2 = max_d2;
  P.n_points = (int)grid.n_points;
  NnGrid& gm = const_cast<NnGrid&>(grid);  // per-grid query scratch
  gm.unresolved.ensure(n + 1);
  gm.unresolved_count.ensure(1);
  B200_CUDA(cudaMemsetAsync(gm.unresolved_count.ptr, 0, sizeof(unsigned), s));
  const int blocks = (int)((n + 127) / 128);
  nn1_kernel<<<blocks, 128, 0, s>>>(P, queries, n, d_idx, d_d2, gm.unresolved_count.ptr, gm.unresolved.ptr);
  nn1_bruteforce_kernel<<<148 * 4, 256, 0, s>>>(P, queries, gm.unresolved_count.ptr, gm.unresolved.ptr, d_idx, d_d2);
  B200_CUDA(cudaGetLastError());
}

void fitness_reduce(const float* d_d2, const int* d_idx, size_t n, double max_range, double* d_scratch2, double* sum,
                    long long* count, cudaStream_t s) {
  B200_CUDA(cudaMemsetAsync(d_scratch2, 0, 2 * sizeof(double), s));
  int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 4);
  if (blocks < 1) blocks = 1;
  fitness_kernel<<<blocks, 256, 0, s>>>(d_d2, d_idx, n, max_range, d_scratch2);
  double res[2];
  B200_CUDA(cudaMemcpyAsync(res, d_scratch2, sizeof(res), cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  *sum = res[0];
  *count = (long long)(res[1] + 0.5);
}

}  // namespace b200
