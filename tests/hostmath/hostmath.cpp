// TEST INFRASTRUCTURE: compiles the __host__ __device__ scalar control math of the solver (csrc/ndt_math.cuh)
// with g++ so that tests/test_hostmath.py can check it on the CPU. Not a compute fallback: nothing here touches
// point clouds.
#include "../../lidarslam_ros2_b200/csrc/ndt_math.cuh"

extern "C" {
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
