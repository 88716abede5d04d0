// TEST INFRASTRUCTURE: compiles the __host__ __device__ scalar control math of the solver (csrc/ndt_math.cuh)
// with g++ so that tests/test_hostmath.py can check it on the CPU. Not a compute fallback: nothing here touches
// point clouds.
#include "../../lidarslam_ros2_b200/csrc/ndt_math.cuh"
#include "../../lidarslam_ros2_b200/csrc/bfgs6.hpp"
#include "../../oracle/bfgs.hpp"

namespace {
// smooth non-quadratic test objective for the BFGS restatements (6 unknowns)
struct TestObjective {
  double c[6], w[6];
  double f(const double* x) const {
    double s = 0;
    for (int i = 0; i < 6; i++) s += w[i] * (x[i] - c[i]) * (x[i] - c[i]);
    return s + 0.1 * (x[0] * x[1]) * (x[0] * x[1]) + 0.5 * (1.0 - cos(x[3] - x[4]));
  }
  void df(const double* x, double* g) const {
    for (int i = 0; i < 6; i++) g[i] = 2 * w[i] * (x[i] - c[i]);
    g[0] += 0.2 * x[0] * x[1] * x[1];
    g[1] += 0.2 * x[0] * x[0] * x[1];
    g[3] += 0.5 * sin(x[3] - x[4]);
    g[4] -= 0.5 * sin(x[3] - x[4]);
  }
  void fdf(const double* x, double& fo, double* g) const {
    fo = f(x);
    df(x, g);
  }
};
}  // namespace

extern "C" {
// runs the product's templated BFGS (plain functor, the form the device instantiation uses) and the oracle's BFGS on the same
// objective; out: x_product[6], x_oracle[6], {f, n_f, n_df, n_fdf, inner} x 2
void hm_bfgs_compare(const double* c6, const double* w6, const double* x0, int max_inner, double grad_tol, double* out22) {
  TestObjective obj;
  for (int i = 0; i < 6; i++) { obj.c[i] = c6[i]; obj.w[i] = w6[i]; }
  {
    b200::Bfgs6T<TestObjective> b(obj);
    double x[6];
    for (int i = 0; i < 6; i++) x[i] = x0[i];
    int inner = 0, result = b.minimizeInit(x);
    result = b200::BFGS_Running;
    do {
      inner++;
      result = b.minimizeOneStep(x);
      if (result) break;
      result = b.testGradient(grad_tol);
    } while (result == b200::BFGS_Running && inner < max_inner);
    for (int i = 0; i < 6; i++) out22[i] = x[i];
    out22[12] = b.f; out22[13] = b.n_f; out22[14] = b.n_df; out22[15] = b.n_fdf; out22[16] = inner;
  }
  {
    oracle::BfgsFunctor6 fn;
    fn.f = [&](const double* x) { return obj.f(x); };
    fn.df = [&](const double* x, double* g) { obj.df(x, g); };
    fn.fdf = [&](const double* x, double& f, double* g) { obj.fdf(x, f, g); };
    oracle::BFGS6 b(fn);
    double x[6];
    for (int i = 0; i < 6; i++) x[i] = x0[i];
    int inner = 0, result = b.minimizeInit(x);
    result = oracle::BFGS_Running;
    do {
      inner++;
      result = b.minimizeOneStep(x);
      if (result) break;
      result = b.testGradient(grad_tol);
    } while (result == oracle::BFGS_Running && inner < max_inner);
    for (int i = 0; i < 6; i++) out22[6 + i] = x[i];
    out22[17] = b.f; out22[18] = b.n_f; out22[19] = b.n_df; out22[20] = b.n_fdf; out22[21] = inner;
  }
}

void hm_angle_tables(const double* p6, float* jang24, float* hang45, double* jd24, double* hd45) {
  b200::angle_tables(p6, jang24, hang45, jd24, hd45);
}
void hm_angle_tables_coded(const double* p6, double* out69) {
  double f[8];
  for (int a = 0; a < 3; a++) {
    double ang = p6[3 + a];
    if (fabs(ang) < 10e-5) { f[2 * a] = 0.0; f[2 * a + 1] = 1.0; } else { f[2 * a] = sin(ang); f[2 * a + 1] = cos(ang); }
  }
  f[6] = 1.0;
  f[7] = 0.0;
  for (int e = 0; e < 69; e++) out69[e] = b200::angle_table_entry(b200::kAngleTableCode[e], f);
}
void hm_sincos_compact(double x, double* sc2) { b200::sincos_compact(x, sc2, sc2 + 1); }
void hm_pose_to_matrix(const double* p6, float* T12) { b200::pose_to_matrix(p6, T12); }
void hm_euler_angles_012(const float* R9, float* out3) { b200::euler_angles_012(R9, out3); }
void hm_solve6(const double* H36, const double* b6, double* x6) { b200::solve6(H36, b6, x6); }
int hm_ldlt_solve6(const double* H36, const double* b6, double* x6) {
  double U[6][6], rhs[6], x[6] = {0, 0, 0, 0, 0, 0};
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) U[r][c] = H36[r * 6 + c];
    rhs[r] = b6[r];
  }
  const bool ok = b200::ldlt_solve6_upper(U, rhs, x);
  for (int k = 0; k < 6; k++) x6[k] = x[k];
  return ok ? 1 : 0;
}
void hm_solve6_svd(const double* H36, const double* b6, double* x6) { b200::solve6_svd(H36, b6, x6); }
double hm_mt_trial(const double* v9) { return b200::mt_trial_value(v9[0], v9[1], v9[2], v9[3], v9[4], v9[5], v9[6], v9[7], v9[8]); }
int hm_mt_update(double* v6, const double* t3) {
  return b200::mt_update_interval(v6[0], v6[1], v6[2], v6[3], v6[4], v6[5], t3[0], t3[1], t3[2]) ? 1 : 0;
}
void hm_gauss(double outlier_ratio, float resolution, double* d3) {
  b200::GaussConsts g = b200::gauss_constants(outlier_ratio, resolution);
  d3[0] = g.d1; d3[1] = g.d2; d3[2] = g.d3;
}
}
