// K1 — fused NDT derivative kernel inside a persistent, cooperative, device-resident Newton loop.
//
// Replaces (Thirdparty/ndt_omp_ros2/include/pclomp/ndt_omp_impl.hpp):
//   computeTransformation :80-171, computeDerivatives :179-284, computePointDerivatives :396-438,
//   updateDerivatives :482-535, computeStepLengthMT :756-916 (+ :632-753), and the neighbourhood lookups
//   voxel_grid_covariance_omp_impl.hpp:373-442; pcl::transformPointCloud (ndt_omp_impl.hpp:100,817,862) is fused
//   into the point load.
//
// B200 design (not a translation of the OpenMP loop):
//   * ONE kernel launch per align(): a cooperative grid of persistent CTAs iterates
//       evaluate (all CTAs) -> grid barrier -> controller (last-arriving CTA: 6x6 solve, line-search state
//       machine, next pose + angle tables) -> release -> evaluate ...
//     so the ~5-40 sequential evaluations of one registration cost no launches and no host round trips.
//   * per (point, voxel) pair only the exponential weight e, s = C x' and the 15 sums S += e s, M += e C,
//     Q += e s s^T are formed; the 6-vector gradient and 6x6 Hessian contribution J^T(.)J is applied once per
//     POINT (J, H_E depend on the point only). ~35 FMA per pair + ~140 per point instead of ~600 MAC per pair.
//   * voxel lookup = occupancy-bitmap rank index (common.cuh), staged into shared memory by a TMA bulk copy
//     (cp.async.bulk + mbarrier) once per launch and reused by every evaluation; 48-byte voxel records are read
//     with three 16-byte read-only loads.
//   * reductions: per-thread f32 sums of <= a few points -> f64 halving-butterfly across the warp (31 shuffled
//     doubles instead of 32*5) -> per-CTA partial -> fixed-order f64 sum over CTAs: bitwise deterministic.
//
// Algorithmic HBM bytes per evaluation: N_src*16 + N_src*probes*8 + N_hit*48 + 28*8.
#include <cooperative_groups.h>

#include "ndt_solver.cuh"

namespace b200 {

namespace {

constexpr int SOLVER_THREADS = 256;
constexpr int SOLVER_WARPS = SOLVER_THREADS / 32;
constexpr long long SPIN_TIMEOUT_CYCLES = 4000000000LL;  // ~2 s: device-side watchdog, never reached in normal runs

struct SharedCtl {
  NdtControl c;
  int is_last;
  int abort;
};

// ---- TMA bulk copy helpers (cp.async.bulk → UBLKCP) -----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned phase) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- warp reduction of 32 doubles per lane: after the call lane L holds the warp total of slot L ----------
template <int N, int S>
struct Butterfly {
  static __device__ __forceinline__ void run(double* v, int lane) {
    const bool upper = (lane & S) != 0;
#pragma unroll
    for (int k = 0; k < N / 2; k++) {
      double send = upper ? v[k] : v[k + N / 2];
      double keep = upper ? v[k + N / 2] : v[k];
      double recv = __shfl_xor_sync(0xffffffffu, send, S);
      v[k] = keep + recv;
    }
    Butterfly<N / 2, S / 2>::run(v, lane);
  }
};
template <>
struct Butterfly<1, 0> {
  static __device__ __forceinline__ void run(double*, int) {}
};

// ---- per-thread accumulators of one evaluation --------------------------------------------------------------
struct Accum {
  float g[6];
  float h[21];
  double score;
  int hits;
};

struct PairSums {  // sums over the voxels hit by one point
  float S0, S1, S2;                      // sum e * s,           s = C x'
  float M00, M01, M02, M11, M12, M22;    // sum e * C
  float Q00, Q01, Q02, Q11, Q12, Q22;    // sum e * s s^T
};

// one (point, voxel) pair — updateDerivatives (ndt_omp_impl.hpp:482-535) reduced to its per-pair core
template <bool HESS>
__device__ __forceinline__ void accumulate_pair(const VoxelRecord* __restrict__ rec, float3 xt, const NdtLaunch& L,
                                                float gd2, PairSums& ps, double& score, int& hits) {
  const double2 m01 = __ldg(reinterpret_cast<const double2*>(rec));
  const double2 m2c = __ldg(reinterpret_cast<const double2*>(rec) + 1);
  const float4 cc = __ldg(reinterpret_cast<const float4*>(rec) + 2);
  const float c00 = __int_as_float(__double2loint(m2c.y)), c01 = __int_as_float(__double2hiint(m2c.y));
  const float c02 = cc.x, c11 = cc.y, c12 = cc.z, c22 = cc.w;
  // x' = x_trans - mean in f64, then to f32 (ndt_omp_impl.hpp:259-262, 490)
  const float x0 = (float)((double)xt.x - m01.x);
  const float x1 = (float)((double)xt.y - m01.y);
  const float x2 = (float)((double)xt.z - m2c.x);
  const float s0 = c00 * x0 + c01 * x1 + c02 * x2;
  const float s1 = c01 * x0 + c11 * x1 + c12 * x2;
  const float s2 = c02 * x0 + c12 * x1 + c22 * x2;
  const float q = x0 * s0 + x1 * s1 + x2 * s2;
  float e = expf(-gd2 * q * 0.5f);                 // :497
  const float score_inc = (float)(-L.d1 * (double)e);  // :499
  e = gd2 * e;                                     // :501
  if (e > 1.0f || e < 0.0f || e != e) return;      // :504-505 (the score increment is dropped too)
  e = (float)((double)e * L.d1);                   // :508
  score += (double)score_inc;
  hits += 1;
  ps.S0 += e * s0;
  ps.S1 += e * s1;
  ps.S2 += e * s2;
  if (HESS) {
    ps.M00 += e * c00; ps.M01 += e * c01; ps.M02 += e * c02;
    ps.M11 += e * c11; ps.M12 += e * c12; ps.M22 += e * c22;
    const float es0 = e * s0, es1 = e * s1, es2 = e * s2;
    ps.Q00 += es0 * s0; ps.Q01 += es0 * s1; ps.Q02 += es0 * s2;
    ps.Q11 += es1 * s1; ps.Q12 += es1 * s2; ps.Q22 += es2 * s2;
  }
}

// per-POINT application of J (point gradient, :396-412) and H_E (:414-436) to the pair sums
template <bool HESS>
__device__ __forceinline__ void apply_point(const float4 p, const PairSums& ps, const SharedCtl& sc, float gd2, Accum& a) {
  const float* ja = sc.c.jang;
  const float x = p.x, y = p.y, z = p.z;
  // J columns 3..5: J3 = (0, j0, j1), J4 = (j2, j3, j4), J5 = (j5, j6, j7)
  const float j0 = ja[0] * x + ja[1] * y + ja[2] * z;
  const float j1 = ja[3] * x + ja[4] * y + ja[5] * z;
  const float j2 = ja[6] * x + ja[7] * y + ja[8] * z;
  const float j3 = ja[9] * x + ja[10] * y + ja[11] * z;
  const float j4 = ja[12] * x + ja[13] * y + ja[14] * z;
  const float j5 = ja[15] * x + ja[16] * y + ja[17] * z;
  const float j6 = ja[18] * x + ja[19] * y + ja[20] * z;
  const float j7 = ja[21] * x + ja[22] * y + ja[23] * z;
  a.g[0] += ps.S0;
  a.g[1] += ps.S1;
  a.g[2] += ps.S2;
  a.g[3] += j0 * ps.S1 + j1 * ps.S2;
  a.g[4] += j2 * ps.S0 + j3 * ps.S1 + j4 * ps.S2;
  a.g[5] += j5 * ps.S0 + j6 * ps.S1 + j7 * ps.S2;
  if (HESS) {
    const float* ha = sc.c.hang;
    // W = sum e (C - d2 s s^T)
    const float W00 = ps.M00 - gd2 * ps.Q00, W01 = ps.M01 - gd2 * ps.Q01, W02 = ps.M02 - gd2 * ps.Q02;
    const float W11 = ps.M11 - gd2 * ps.Q11, W12 = ps.M12 - gd2 * ps.Q12, W22 = ps.M22 - gd2 * ps.Q22;
    // W * J3, W * J4, W * J5
    const float a0 = W01 * j0 + W02 * j1, a1 = W11 * j0 + W12 * j1, a2 = W12 * j0 + W22 * j1;
    const float b0 = W00 * j2 + W01 * j3 + W02 * j4, b1 = W01 * j2 + W11 * j3 + W12 * j4, b2 = W02 * j2 + W12 * j3 + W22 * j4;
    const float c0 = W00 * j5 + W01 * j6 + W02 * j7, c1 = W01 * j5 + W11 * j6 + W12 * j7, c2 = W02 * j5 + W12 * j6 + W22 * j7;
    // second-derivative vectors a..f dotted with S (rows of hang: a2 a3 b2 b3 c2 c3 d1 d2 d3 e1 e2 e3 f1 f2 f3)
    const float hA2 = ha[0] * x + ha[1] * y + ha[2] * z, hA3 = ha[3] * x + ha[4] * y + ha[5] * z;
    const float hB2 = ha[6] * x + ha[7] * y + ha[8] * z, hB3 = ha[9] * x + ha[10] * y + ha[11] * z;
    const float hC2 = ha[12] * x + ha[13] * y + ha[14] * z, hC3 = ha[15] * x + ha[16] * y + ha[17] * z;
    const float hD1 = ha[18] * x + ha[19] * y + ha[20] * z, hD2 = ha[21] * x + ha[22] * y + ha[23] * z,
                hD3 = ha[24] * x + ha[25] * y + ha[26] * z;
    const float hE1 = ha[27] * x + ha[28] * y + ha[29] * z, hE2 = ha[30] * x + ha[31] * y + ha[32] * z,
                hE3 = ha[33] * x + ha[34] * y + ha[35] * z;
    const float hF1 = ha[36] * x + ha[37] * y + ha[38] * z, hF2 = ha[39] * x + ha[40] * y + ha[41] * z,
                hF3 = ha[42] * x + ha[43] * y + ha[44] * z;
    // upper triangle, row-major: (0,0..5) (1,1..5) (2,2..5) (3,3..5) (4,4..5) (5,5)
    a.h[0] += W00; a.h[1] += W01; a.h[2] += W02; a.h[3] += a0; a.h[4] += b0; a.h[5] += c0;
    a.h[6] += W11; a.h[7] += W12; a.h[8] += a1; a.h[9] += b1; a.h[10] += c1;
    a.h[11] += W22; a.h[12] += a2; a.h[13] += b2; a.h[14] += c2;
    a.h[15] += (j0 * a1 + j1 * a2) + (hA2 * ps.S1 + hA3 * ps.S2);
    a.h[16] += (j0 * b1 + j1 * b2) + (hB2 * ps.S1 + hB3 * ps.S2);
    a.h[17] += (j0 * c1 + j1 * c2) + (hC2 * ps.S1 + hC3 * ps.S2);
    a.h[18] += (j2 * b0 + j3 * b1 + j4 * b2) + (hD1 * ps.S0 + hD2 * ps.S1 + hD3 * ps.S2);
    a.h[19] += (j2 * c0 + j3 * c1 + j4 * c2) + (hE1 * ps.S0 + hE2 * ps.S1 + hE3 * ps.S2);
    a.h[20] += (j5 * c0 + j6 * c1 + j7 * c2) + (hF1 * ps.S0 + hF2 * ps.S1 + hF3 * ps.S2);
  }
}

// neighbourhood of one transformed point for the four pclomp::NeighborSearchMethod values
template <int METHOD, bool HESS, bool STAGED>
__device__ __forceinline__ void process_point(const NdtLaunch& L, const SharedCtl& sc, const RankWord* sidx, int i,
                                              float gd2, Accum& acc) {
  const float4 p = L.src[i];
  const float3 xt = transform_point(sc.c.T, p);
  const int ci = lookup_cell(xt.x, L.geom.leaf), cj = lookup_cell(xt.y, L.geom.leaf), ck = lookup_cell(xt.z, L.geom.leaf);
  PairSums ps = {};
  const int hits_before = acc.hits;
  if (METHOD == 2) {  // DIRECT7 (voxel_grid_covariance_omp_impl.hpp:418-433)
    int r[7];
    r[0] = probe_cell<STAGED>(L.geom, L.index, sidx, ci, cj, ck);
    r[1] = probe_cell<STAGED>(L.geom, L.index, sidx, ci + 1, cj, ck);
    r[2] = probe_cell<STAGED>(L.geom, L.index, sidx, ci - 1, cj, ck);
    r[3] = probe_cell<STAGED>(L.geom, L.index, sidx, ci, cj + 1, ck);
    r[4] = probe_cell<STAGED>(L.geom, L.index, sidx, ci, cj - 1, ck);
    r[5] = probe_cell<STAGED>(L.geom, L.index, sidx, ci, cj, ck + 1);
    r[6] = probe_cell<STAGED>(L.geom, L.index, sidx, ci, cj, ck - 1);
#pragma unroll
    for (int k = 0; k < 7; k++)
      if (r[k] >= 0) accumulate_pair<HESS>(L.records + r[k], xt, L, gd2, ps, acc.score, acc.hits);
  } else if (METHOD == 3) {  // DIRECT1
    int r0 = probe_cell<STAGED>(L.geom, L.index, sidx, ci, cj, ck);
    if (r0 >= 0) accumulate_pair<HESS>(L.records + r0, xt, L, gd2, ps, acc.score, acc.hits);
  } else {  // DIRECT26 (26 cells, centre excluded) / KDTREE (27 cells + centroid radius test)
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (METHOD == 1 && dx == 0 && dy == 0 && dz == 0) continue;
          int r = probe_cell<STAGED>(L.geom, L.index, sidx, ci + dx, cj + dy, ck + dz);
          if (r < 0) continue;
          if (METHOD == 0) {  // radiusSearch over voxel centroids (voxel_grid_covariance_omp.h:470-499)
            const float4 c = __ldg(L.centroids + r);
            const float ex = __fsub_rn(xt.x, c.x), ey = __fsub_rn(xt.y, c.y), ez = __fsub_rn(xt.z, c.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
            if (!(d2 < L.radius2)) continue;
          }
          accumulate_pair<HESS>(L.records + r, xt, L, gd2, ps, acc.score, acc.hits);
        }
  }
  if (acc.hits != hits_before) apply_point<HESS>(p, ps, sc, gd2, acc);
}

template <int METHOD, bool STAGED>
__device__ __forceinline__ void evaluate(const NdtLaunch& L, const SharedCtl& sc, const RankWord* sidx, Accum& acc) {
  const float gd2 = (float)L.d2;
  const int stride = gridDim.x * SOLVER_THREADS;
  if (sc.c.compute_hessian) {
    for (int i = blockIdx.x * SOLVER_THREADS + threadIdx.x; i < L.n_src; i += stride)
      process_point<METHOD, true, STAGED>(L, sc, sidx, i, gd2, acc);
  } else {
    for (int i = blockIdx.x * SOLVER_THREADS + threadIdx.x; i < L.n_src; i += stride)
      process_point<METHOD, false, STAGED>(L, sc, sidx, i, gd2, acc);
  }
}

// =====================================================================================================
// controller: runs in ONE thread of the last-arriving CTA after every evaluation
// =====================================================================================================
__device__ void write_control(NdtSolverWork* W, const double* x_t, int compute_hessian, bool want_f64_tables) {
  NdtControl& c = W->control;
  pose_to_matrix(x_t, c.T);
  angle_tables(x_t, c.jang, c.hang, want_f64_tables ? W->state.jd : nullptr, want_f64_tables ? W->state.hd : nullptr);
  c.mode = EVAL_DERIV;
  c.compute_hessian = compute_hessian;
  float* F = W->state.final_T;  // final_transformation_ (ndt_omp_impl.hpp:811-814)
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 4; k++) F[r * 4 + k] = c.T[r * 4 + k];
  F[12] = F[13] = F[14] = 0.0f;
  F[15] = 1.0f;
}

__device__ void load_totals(NdtState& st, const double* tot, bool with_hessian) {
  st.score = tot[SLOT_SCORE];
  for (int k = 0; k < 6; k++) st.g[k] = tot[SLOT_G + k];
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) {
      double v = with_hessian ? tot[SLOT_H + tri_index(i, j)] : 0.0;
      st.H[i * 6 + j] = v;
      st.H[j * 6 + i] = v;
    }
  st.hits_last = (long long)(tot[SLOT_HITS] + 0.5);
  st.hits_total += st.hits_last;
  st.evaluations += 1;
}

__device__ void finish(const NdtLaunch& L, NdtSolverWork* W) {
  NdtState& st = W->state;
  NdtResult& r = W->result;
  for (int k = 0; k < 16; k++) r.final_T[k] = st.final_T[k];
  r.score = st.score;
  r.trans_probability = st.score / (double)L.n_src;  // ndt_omp_impl.hpp:136,170
  for (int k = 0; k < 6; k++) r.g[k] = st.g[k];
  for (int k = 0; k < 36; k++) r.H[k] = st.H[k];
  r.hits_last = st.hits_last;
  r.hits_total = st.hits_total;
  r.converged = st.converged;
  r.iterations = st.nr_iterations;
  r.evaluations = st.evaluations;
  r.error = 0;
  W->control.mode = EVAL_DONE;
}

__device__ __noinline__ void controller(const NdtLaunch& L, NdtSolverWork* W, const double* tot) {
  NdtState& st = W->state;
  const double mu = 1.e-4, nu = 0.9;  // ndt_omp_impl.hpp:788-790
  const double step_max = L.step_size, step_min = L.trans_eps / 2;
  enum { ACT_NEWTON_BEGIN, ACT_NEWTON_END, ACT_LS_CHECK, ACT_RETURN };
  int act;
  double phi_t = 0, d_phi_t = 0, psi_t = 0, d_psi_t = 0;

  if (L.mode != NDT_MODE_ALIGN) {  // single derivative pass requested through the C-ABI
    load_totals(st, tot, L.init.compute_hessian != 0);
    st.converged = 0;
    finish(L, W);
    return;
  }

  auto eval_point_values = [&]() {
    phi_t = -st.score;
    double dd = 0;
    for (int k = 0; k < 6; k++) dd += st.g[k] * st.dir[k];
    d_phi_t = -dd;
    psi_t = mt_psi(st.a_t, phi_t, st.phi_0, st.d_phi_0, mu);
    d_psi_t = mt_dpsi(d_phi_t, st.d_phi_0, mu);
  };

  switch (st.phase) {
    case PH_INITIAL:  // result of the initial computeDerivatives (:119)
      load_totals(st, tot, true);
      act = ACT_NEWTON_BEGIN;
      break;
    case PH_LS_FIRST:  // first evaluation inside computeStepLengthMT (:821)
      load_totals(st, tot, true);
      eval_point_values();
      act = ACT_LS_CHECK;
      break;
    case PH_LS_ITER:  // More-Thuente inner evaluation (:865), compute_hessian = false zeroes the Hessian
      load_totals(st, tot, false);
      eval_point_values();
      if (st.open_interval && (psi_t <= 0 && d_psi_t >= 0)) {  // :878-889
        st.open_interval = 0;
        st.f_l = st.f_l + st.phi_0 - mu * st.d_phi_0 * st.a_l;
        st.g_l = st.g_l + mu * st.d_phi_0;
        st.f_u = st.f_u + st.phi_0 - mu * st.d_phi_0 * st.a_u;
        st.g_u = st.g_u + mu * st.d_phi_0;
      }
      if (st.open_interval)
        st.interval_converged = mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, psi_t, d_psi_t);
      else
        st.interval_converged = mt_update_interval(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, phi_t, d_phi_t);
      st.step_iterations++;
      act = ACT_LS_CHECK;
      break;
    default:  // PH_LS_HESSIAN: the K2 pass has written st.H (:912-913)
      act = ACT_NEWTON_END;
      break;
  }

  for (int guard = 0; guard < 8; guard++) {
    if (act == ACT_LS_CHECK) {
      // :834
      if (!st.interval_converged && st.step_iterations < 10 && !(psi_t <= 0 && d_phi_t <= -nu * st.d_phi_0)) {
        if (st.open_interval)
          st.a_t = mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, psi_t, d_psi_t);
        else
          st.a_t = mt_trial_value(st.a_l, st.f_l, st.g_l, st.a_u, st.f_u, st.g_u, st.a_t, phi_t, d_phi_t);
        st.a_t = fmin(st.a_t, step_max);
        st.a_t = fmax(st.a_t, step_min);
        for (int k = 0; k < 6; k++) st.x_t[k] = st.p[k] + st.dir[k] * st.a_t;
        write_control(W, st.x_t, 0, true);
        st.phase = PH_LS_ITER;
        return;
      }
      if (st.step_iterations) {  // :912-913 — needs the f64 radius-neighbourhood Hessian (K2): leave the kernel
        st.phase = PH_LS_HESSIAN;
        W->control.mode = EVAL_NEED_HESSIAN;
        W->result.error = 100;  // host: run the K2 pass, then resume
        return;
      }
      act = ACT_NEWTON_END;
    }
    if (act == ACT_NEWTON_END) {
      // :143-164
      for (int k = 0; k < 6; k++) st.p[k] = st.p[k] + st.dir[k] * st.a_t;
      if (st.nr_iterations > L.max_iterations || (st.nr_iterations && (fabs(st.a_t) < L.trans_eps))) st.converged = 1;
      st.nr_iterations++;
      if (st.converged) {
        finish(L, W);
        return;
      }
      act = ACT_NEWTON_BEGIN;
    }
    if (act == ACT_NEWTON_BEGIN) {
      // :127-142 and the prologue of computeStepLengthMT :761-821
      double neg_g[6], dp[6];
      for (int k = 0; k < 6; k++) neg_g[k] = -st.g[k];
      solve6(st.H, neg_g, dp);
      double n2 = 0;
      for (int k = 0; k < 6; k++) n2 += dp[k] * dp[k];
      const double norm = sqrt(n2);
      if (norm == 0 || norm != norm) {
        st.converged = (norm == norm) ? 1 : 0;
        finish(L, W);
        return;
      }
      for (int k = 0; k < 6; k++) st.dir[k] = dp[k] / norm;
      st.phi_0 = -st.score;
      double dd = 0;
      for (int k = 0; k < 6; k++) dd += st.g[k] * st.dir[k];
      st.d_phi_0 = -dd;
      if (st.d_phi_0 >= 0) {
        if (st.d_phi_0 == 0) {  // :771-772: zero step, no evaluation
          st.a_t = 0;
          act = ACT_NEWTON_END;
          continue;
        }
        st.d_phi_0 *= -1;
        for (int k = 0; k < 6; k++) st.dir[k] *= -1;
      }
      st.step_iterations = 0;
      st.a_l = 0;
      st.a_u = 0;
      st.f_l = mt_psi(st.a_l, st.phi_0, st.phi_0, st.d_phi_0, mu);
      st.g_l = mt_dpsi(st.d_phi_0, st.d_phi_0, mu);
      st.f_u = mt_psi(st.a_u, st.phi_0, st.phi_0, st.d_phi_0, mu);
      st.g_u = mt_dpsi(st.d_phi_0, st.d_phi_0, mu);
      st.interval_converged = (step_max - step_min) > 0 ? 1 : 0;  // :803 (sic)
      st.open_interval = 1;
      st.a_t = norm;
      st.a_t = fmin(st.a_t, step_max);
      st.a_t = fmax(st.a_t, step_min);
      for (int k = 0; k < 6; k++) st.x_t[k] = st.p[k] + st.dir[k] * st.a_t;
      write_control(W, st.x_t, 1, !st.interval_converged);
      st.phase = PH_LS_FIRST;
      return;
    }
  }
  // unreachable in practice (two consecutive zero-step iterations terminate); fail safe
  st.converged = 0;
  finish(L, W);
}

// =====================================================================================================
// the persistent kernel
// =====================================================================================================
template <int METHOD>
__global__ void __launch_bounds__(SOLVER_THREADS, 2) ndt_solver_kernel(const __grid_constant__ NdtLaunch L) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ SharedCtl sc;
  __shared__ double warp_part[SOLVER_WARPS][SLOT_COUNT];
  __shared__ double tot[SLOT_COUNT];
  __shared__ __align__(8) unsigned long long tma_bar;

  NdtSolverWork* W = L.work;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const RankWord* sidx = reinterpret_cast<const RankWord*>(dyn_smem);

  // ---- stage the voxel rank index into shared memory with TMA bulk copies (once per launch) -------------
  if (L.index_in_smem) {
    const unsigned bytes = ((unsigned)L.geom.n_words * 8u + 15u) & ~15u;
    if (tid == 0) {
      mbar_init(&tma_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&tma_bar, bytes);
      for (unsigned off = 0; off < bytes; off += 16384u) {
        unsigned chunk = min(16384u, bytes - off);
        tma_bulk_g2s(dyn_smem + off, reinterpret_cast<const unsigned char*>(L.index) + off, chunk, &tma_bar);
      }
    }
    long long t0 = clock64();
    while (!mbar_try_wait(&tma_bar, 0)) {
      if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) break;
    }
  }

  unsigned my_gen = 0;
  if (tid == 0) {
    my_gen = ld_relaxed_gpu(&W->gen);
    sc.abort = 0;
    sc.is_last = 0;
  }
  // round-0 control: from the launch parameters (fresh solve) or from the work area (resume after K2)
  {
    const int* src = L.resume ? reinterpret_cast<const int*>(&W->control) : reinterpret_cast<const int*>(&L.init);
    int* dst = reinterpret_cast<int*>(&sc.c);
    for (int k = tid; k < NDT_CONTROL_WORDS; k += SOLVER_THREADS) dst[k] = L.resume ? __ldcg(src + k) : src[k];
  }
  if (!L.resume && blockIdx.x == 0 && tid == 0) {
    // fresh controller state; only the CTA that later runs the controller reads it, after a grid barrier
    NdtState& st = W->state;
    for (int k = 0; k < 6; k++) st.p[k] = L.p0[k];
    for (int k = 0; k < 16; k++) st.final_T[k] = L.init_final[k];
    st.phase = PH_INITIAL;
    st.nr_iterations = 0;
    st.evaluations = 0;
    st.converged = 0;
    st.hits_total = 0;
    st.hits_last = 0;
    st.step_iterations = 0;
    st.a_t = 0;
    W->result.error = 2;  // "not finished"; finish() sets 0, the watchdog 1, a K2 request 100
  }
  bool skip_eval = L.resume != 0;

  for (;;) {
    __syncthreads();
    if (sc.c.mode != EVAL_DERIV || sc.abort) break;

    // ---- (1) evaluate this CTA's points ---------------------------------------------------------------
    Accum acc;
#pragma unroll
    for (int k = 0; k < 6; k++) acc.g[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 21; k++) acc.h[k] = 0.f;
    acc.score = 0.0;
    acc.hits = 0;
    if (!skip_eval) {
      if (L.index_in_smem) evaluate<METHOD, true>(L, sc, sidx, acc);
      else evaluate<METHOD, false>(L, sc, sidx, acc);
    }
    skip_eval = false;

    // ---- (2) warp butterfly in f64, CTA partial --------------------------------------------------------
    {
      double v[SLOT_COUNT];
      v[SLOT_SCORE] = acc.score;
#pragma unroll
      for (int k = 0; k < 6; k++) v[SLOT_G + k] = (double)acc.g[k];
#pragma unroll
      for (int k = 0; k < 21; k++) v[SLOT_H + k] = (double)acc.h[k];
      v[SLOT_HITS] = (double)acc.hits;
      v[29] = 0.0;
      v[30] = 0.0;
      v[31] = 0.0;
      Butterfly<32, 16>::run(v, lane);
      warp_part[warp][lane] = v[0];
    }
    __syncthreads();
    if (tid < SLOT_COUNT) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < SOLVER_WARPS; w++) s += warp_part[w][tid];
      W->partials[blockIdx.x][tid] = s;
    }
    __syncthreads();

    // ---- (3) grid barrier; the last CTA to arrive reduces and runs the controller -----------------------
    if (tid == 0) {
      __threadfence();
      unsigned prev = atom_add_acq_rel_gpu(&W->arrive, 1u);
      sc.is_last = (prev == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sc.is_last) {
      __threadfence();
      double s = 0;
      for (int r = warp; r < (int)gridDim.x; r += SOLVER_WARPS) s += __ldcg(&W->partials[r][lane]);
      warp_part[warp][lane] = s;
      __syncthreads();
      if (tid < SLOT_COUNT) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < SOLVER_WARPS; w++) t += warp_part[w][tid];
        tot[tid] = t;
      }
      __syncthreads();
      if (tid == 0) {
        controller(L, W, tot);
        W->arrive = 0;
        __threadfence();
        st_release_gpu(&W->gen, my_gen + 1);
      }
    } else if (tid == 0) {
      long long t0 = clock64();
      while (ld_relaxed_gpu(&W->gen) == my_gen) {
        if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) {
          atomicExch(&W->error, 1u);
          W->result.error = 1;
          sc.abort = 1;
          break;
        }
      }
      fence_acq_rel_gpu();
    }
    if (tid == 0) my_gen += 1;
    __syncthreads();
    // ---- (4) next round's control block (written by the controller; read through L2) --------------------
    {
      const int* src = reinterpret_cast<const int*>(&W->control);
      int* dst = reinterpret_cast<int*>(&sc.c);
      for (int k = tid; k < NDT_CONTROL_WORDS; k += SOLVER_THREADS) dst[k] = __ldcg(src + k);
    }
  }
}

using KernelFn = void (*)(const NdtLaunch);
KernelFn kernel_for(int method) {
  switch (method) {
    case 0: return ndt_solver_kernel<0>;
    case 1: return ndt_solver_kernel<1>;
    case 3: return ndt_solver_kernel<3>;
    default: return ndt_solver_kernel<2>;
  }
}

}  // namespace

// =====================================================================================================
// host side
// =====================================================================================================
NdtSolver::~NdtSolver() {
  if (d_work_) cudaFree(d_work_);
  if (h_result_) cudaFreeHost(h_result_);
}

void NdtSolver::init(int device, cudaStream_t s) {
  device_ = device;
  stream_ = s;
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  sm_count_ = prop.multiProcessorCount;
  max_smem_optin_ = (int)prop.sharedMemPerBlockOptin;
  B200_CUDA(cudaMalloc(&d_work_, sizeof(NdtSolverWork)));
  B200_CUDA(cudaMemset(d_work_, 0, sizeof(NdtSolverWork)));
  B200_CUDA(cudaMallocHost(&h_result_, sizeof(NdtResult)));
  std::memset(h_result_, 0, sizeof(NdtResult));
  for (int m = 0; m < 4; m++)
    B200_CUDA(cudaFuncSetAttribute(kernel_for(m), cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
}

const double* NdtSolver::state_jd() const { return d_work_->state.jd; }
const double* NdtSolver::state_hd() const { return d_work_->state.hd; }
const float* NdtSolver::control_T() const { return d_work_->control.T; }

void NdtSolver::reset_barrier() {
  B200_CUDA(cudaMemsetAsync(d_work_, 0, 16, stream_));  // arrive, gen, error, pad
}

void NdtSolver::launch(const VoxelMap& map, const float4* src, size_t n_src, const NdtConfig& cfg, int mode,
                       const float* T_rowmajor16, const double* p6, int compute_hessian, int resume) {
  NdtLaunch L{};
  L.src = src;
  L.index = map.index.ptr;
  L.records = map.records.ptr;
  L.icov_d = map.icov_d.ptr;
  L.centroids = map.centroids.ptr;
  L.work = d_work_;
  L.geom = map.geom;
  L.n_src = (int)n_src;
  L.n_voxels = (int)map.n_voxels;
  L.search_method = cfg.search_method;
  L.mode = mode;
  L.resume = resume;
  L.max_iterations = cfg.max_iterations;
  L.resolution = cfg.resolution;
  L.radius2 = static_cast<float>((double)cfg.resolution * (double)cfg.resolution);
  GaussConsts gc = gauss_constants(cfg.outlier_ratio, cfg.resolution);
  L.d1 = gc.d1;
  L.d2 = gc.d2;
  L.d3 = gc.d3;
  L.step_size = cfg.step_size;
  L.trans_eps = cfg.trans_eps;

  // initial pose: final_transformation_ = guess (or identity), p = (t, eulerAngles(0,1,2)) in float → double
  // (ndt_omp_impl.hpp:95-111); the first evaluation transforms the source by the guess matrix itself.
  float T[16];
  std::memcpy(T, T_rowmajor16, sizeof(T));
  std::memcpy(L.init_final, T, sizeof(T));
  for (int k = 0; k < 12; k++) L.init.T[k] = T[k];
  double p0[6];
  if (p6) {
    for (int k = 0; k < 6; k++) p0[k] = p6[k];
  } else {
    float R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    float ang[3];
    euler_angles_012(R, ang);
    p0[0] = T[3];
    p0[1] = T[7];
    p0[2] = T[11];
    p0[3] = ang[0];
    p0[4] = ang[1];
    p0[5] = ang[2];
  }
  for (int k = 0; k < 6; k++) L.p0[k] = p0[k];
  angle_tables(p0, L.init.jang, L.init.hang, nullptr, nullptr);
  L.init.mode = EVAL_DERIV;
  L.init.compute_hessian = compute_hessian;

  // shared-memory staging of the rank index when it fits
  const size_t index_bytes = ((size_t)map.geom.n_words * 8 + 15) & ~(size_t)15;
  L.index_in_smem = (map.geom.n_words > 0 && index_bytes <= 64 * 1024) ? 1 : 0;
  const size_t dyn_smem = L.index_in_smem ? index_bytes : 0;

  KernelFn fn = kernel_for(cfg.search_method);
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, SOLVER_THREADS, dyn_smem));
  if (per_sm < 1) throw CudaError("ndt_solver_kernel does not fit on an SM");
  int max_ctas = std::min(per_sm * sm_count_, NDT_MAX_CTAS);
  int want = (int)((n_src + SOLVER_THREADS - 1) / SOLVER_THREADS);
  grid_ = std::max(1, std::min(max_ctas, want));
  block_ = SOLVER_THREADS;
  index_in_smem_ = L.index_in_smem;

  void* args[] = {&L};
  B200_CUDA(cudaLaunchCooperativeKernel((const void*)fn, dim3(grid_), dim3(SOLVER_THREADS), args, dyn_smem, stream_));
  launches += 1;
  B200_CUDA(cudaMemcpyAsync(h_result_, &d_work_->result, sizeof(NdtResult), cudaMemcpyDeviceToHost, stream_));
}

}  // namespace b200
