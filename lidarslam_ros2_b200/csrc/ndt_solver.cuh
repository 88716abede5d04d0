// Device-side structures shared by the NDT solver kernels (ndt_solver.cu, ndt_aux.cu).
#pragma once
#include "common.cuh"
#include "engine.hpp"

namespace b200 {

constexpr int NDT_MAX_CTAS = 256;  // one CTA per SM
// Registrations in flight inside one batch launch = controller CTAs. Every launch (single or batched) leaves this many
// SMs to controllers, so that the evaluator count — and with it the point partition and the fixed summation order —
// is the same for b200reg_align and for b200reg_ndt_align_batch: a batched result is bitwise the single-align result.
#ifndef B200_NDT_MAX_SLOTS
#define B200_NDT_MAX_SLOTS 3  // developer switch for A/B builds
#endif
constexpr int NDT_MAX_SLOTS = B200_NDT_MAX_SLOTS;

enum EvalMode : int {
  EVAL_DERIV = 0,        // fused derivative pass (K1)
  EVAL_DONE = 1,         // solver finished, result written
  EVAL_NEED_HESSIAN = 2  // leave the persistent kernel: host runs the f64 radius-Hessian pass (K2) and resumes
};

enum SolverPhase : int { PH_INITIAL = 0, PH_LS_FIRST = 1, PH_LS_ITER = 2, PH_LS_HESSIAN = 3 };

// what every CTA needs for one evaluation round; written by the controller (or by the host for round 0)
struct alignas(16) NdtControl {  // size is a multiple of 16 bytes: arrays of it are read with 16-byte shared-memory loads
  float T[12];     // 3x4 row-major transform applied to the source points
  float jang[24];  // 8 x 3 f32 angle-Jacobian table   (ndt_omp_impl.hpp:337-345)
  float hang[45];  // 15 x 3 f32 angle-Hessian table    (ndt_omp_impl.hpp:371-391)
  int mode;        // EvalMode
  int compute_hessian;
  int job;         // batch launches: index of the registration this block belongs to (evaluators restage on a change)
  int pad[4];
};
static_assert(sizeof(NdtControl) % 16 == 0, "NdtControl must keep 16-byte alignment in arrays");
constexpr int NDT_CONTROL_WORDS = sizeof(NdtControl) / 4;

// controller state (lives in global memory: a different CTA may run the controller every round)
struct NdtState {
  double p[6], score, g[6], H[36];
  double dir[6], x_t[6];
  double jd[24], hd[45];  // f64 angle tables at x_t, consumed by the K2 pass
  double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t;
  float final_T[16];
  long long hits_last, hits_total;
  int interval_converged, open_interval, step_iterations;
  int phase, nr_iterations, evaluations, converged;
};

constexpr int NDT_TIMING_ROUNDS = 48;
constexpr int NDT_TIMING_SLOTS = 12;
// slots (globaltimer ns): 0 CTA0 round start, 1 CTA0 after evaluate, 2/3 CTA0 partial row stored,
//                         4 last CTA detected, 5 partials reduced, 6 controller done, 7 CTA0 released
// Signalling between the evaluator CTAs and the controller CTA carries its own validity ("flag in data", as in NCCL's
// LL protocol), so neither direction needs a counter, a fence or a second dependent round trip:
//  * control block, controller -> evaluators: every 32-bit word travels as one 64-bit store {payload, sequence}; an
//    evaluator thread polls ITS word until the sequence number is the expected one. NDT_CTL_COPIES replicas spread
//    the pollers over L2 lines.
//  * partial sums, evaluators -> controller: rows of 32 doubles, double-buffered by round parity; an unwritten slot
//    holds NDT_PARTIAL_EMPTY (a NaN payload no computation produces). The controller's reduction loads double as the
//    poll; it re-arms each row after consuming it, two rounds before the row is written again.
constexpr int NDT_CTL_COPIES = 4;
constexpr int NDT_CTL_LL_WORDS = 96;  // >= NDT_CONTROL_WORDS, whole 128-byte lines
constexpr unsigned long long NDT_PARTIAL_EMPTY = 0xFFF8DEADFFF8DEADull;
constexpr int NDT_MAX_ROUNDS = 60000;  // sequence numbers are epoch * 65536 + round + 1

// one registration of a batch launch (device array, filled by the host before the launch)
struct NdtJob {
  // source points: records of `stride` bytes with x, y, z floats first — 16 for a float4 cloud resident in HBM, the
  // caller's record size when the raw host records were copied straight in (no unpack pass: the evaluators read them
  // once, when they stage the registration's points into shared memory)
  const unsigned char* src;
  int n_src;
  int stride;
  // ready == nullptr: the points are there when the launch starts. Otherwise the controller waits, before it starts this
  // registration, until *ready == ready_tag: the copy stream writes the tag behind the scan's DMA (cuStreamWriteValue32),
  // so later scans of a batch are still being uploaded while earlier ones are already being registered.
  const unsigned* ready;
  unsigned ready_tag;
  int pad;
  double p0[6];         // initial pose parameters (ndt_omp_impl.hpp:103-111)
  float init_final[16]; // final_transformation_ = guess
  NdtControl init;      // control block of the first evaluation (transform = guess, angle tables at p0)
};

struct NdtSolverWork {
  unsigned error;
  unsigned next_job;     // batch launches: next unassigned registration (atomicAdd by the controllers)
  unsigned pad[2];
  NdtControl control;    // plain copy of the control block, written only when the kernel leaves for a K2 pass
  NdtState state;
  NdtResult result;
  alignas(128) unsigned long long ctl_ll[NDT_MAX_SLOTS][NDT_CTL_COPIES][NDT_CTL_LL_WORDS];
  alignas(128) double partials[NDT_MAX_SLOTS][2][NDT_MAX_CTAS][SLOT_COUNT];
  unsigned long long timing[NDT_TIMING_ROUNDS][NDT_TIMING_SLOTS];
  unsigned cta_eval_ns[NDT_MAX_CTAS][4];  // timing mode, round 2, low 32 bits of globaltimer: start, evaluate end, published
};

struct NdtLaunch {
  const float4* src;
  const RankWord* index;
  const VoxelRecord* records;
  const double* icov_d;
  const float4* centroids;
  NdtSolverWork* work;
  NdtResult* result_host;  // pinned, device-visible host memory: the controller CTA writes the result there on exit
  // batch launches (n_slots >= 1 and jobs != nullptr): n_jobs registrations against the same map, n_slots in flight;
  // result_host is then an array of n_jobs results. jobs == nullptr: the single registration described inline below.
  const NdtJob* jobs;
  int n_jobs;
  int n_slots;
  PoseBoardView board;  // board.world > 0: finished poses are also stored into every peer's pose board (engine.hpp)
  GridGeom geom;
  int n_src;
  int n_voxels;
  int search_method;
  int mode;    // NdtMode
  unsigned epoch;         // launch counter of this handle (high half of the control block's sequence numbers)
  int acc_offset;         // byte offset of the per-thread accumulators in dynamic shared memory (after the rank index)
  int pts_offset;         // byte offset of the staged source points (n_slots x SMEM_POINTS float4) after the accumulators
  int scalar_controller;  // 1: disable the warp-parallel controller fast path (developer switch)
  int timing;  // 1: record per-phase globaltimer stamps into work->timing (developer instrumentation)
  int resume;  // 1: state/control already in work (after a K2 pass); first round skips the evaluation
  int index_in_smem;
  int max_iterations;
  float resolution;
  float radius2;  // (float)(resolution * resolution) in double, the FLANN radius of KDTREE mode
  double d1, d2, d3;
  double step_size, trans_eps;
  double p0[6];
  float init_final[16];
  NdtControl init;
};

// ---- rank-index probe ----------------------------------------------------------------------------------
// returns the record index of the voxel at absolute cell coordinates (ci, cj, ck), or -1
// (VoxelGridCovariance::getNeighborhoodAtPoint, voxel_grid_covariance_omp_impl.hpp:382-399)
template <bool STAGED>
__device__ __forceinline__ int probe_cell(const GridGeom& g, const RankWord* __restrict__ gidx,
                                          const RankWord* __restrict__ sidx, int ci, int cj, int ck) {
  if (ci < g.min_b[0] || ci > g.max_b[0] || cj < g.min_b[1] || cj > g.max_b[1] || ck < g.min_b[2] || ck > g.max_b[2])
    return -1;
  const int lin = (ci - g.min_b[0]) + (cj - g.min_b[1]) * g.mul[1] + (ck - g.min_b[2]) * g.mul[2];
  uint2 w;
  if (STAGED) w = *reinterpret_cast<const uint2*>(sidx + (lin >> 5));
  else w = __ldg(reinterpret_cast<const uint2*>(gidx + (lin >> 5)));
  const unsigned bit = lin & 31;
  if (!((w.x >> bit) & 1u)) return -1;
  return (int)(w.y + __popc(w.x & ((1u << bit) - 1u)));
}

__device__ __forceinline__ float3 transform_point(const float* T, float4 p) {
  // ((T0*x + T1*y) + T2*z) + T3, un-fused like pcl::transformPointCloud's float arithmetic
  float3 r;
  r.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], p.x), __fmul_rn(T[1], p.y)), __fmul_rn(T[2], p.z)), T[3]);
  r.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], p.x), __fmul_rn(T[5], p.y)), __fmul_rn(T[6], p.z)), T[7]);
  r.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], p.x), __fmul_rn(T[9], p.y)), __fmul_rn(T[10], p.z)), T[11]);
  return r;
}

// lookup cell of a transformed point: floor(x / leaf) with an IEEE division (impl.hpp:379-381)
__device__ __forceinline__ int lookup_cell(float x, float leaf) { return (int)floorf(__fdiv_rn(x, leaf)); }
// Same result without the division on the common path: q = x * (1/leaf) is within 2 ulp of the IEEE quotient, so
// floor(q) can only differ from floor(x / leaf) when q lies within a few ulp of an integer — only then is the exact
// division evaluated.
__device__ __forceinline__ int lookup_cell_fast(float x, float leaf, float inv_leaf) {
  const float q = __fmul_rn(x, inv_leaf);
  if (fabsf(q - rintf(q)) <= 1e-6f * fabsf(q) + 1e-30f) return (int)floorf(__fdiv_rn(x, leaf));
  return (int)floorf(q);
}

}  // namespace b200
