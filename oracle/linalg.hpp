// ORACLE — TEST INFRASTRUCTURE ONLY (parity unpinned: the reference ships no tests).
// Small dense linear algebra the reference obtains from Eigen 3.4 (not vendored, absent here):
//   SelfAdjointEigenSolver<Matrix3d>   (voxel_grid_covariance_omp_impl.hpp:333)
//   Matrix3d::inverse                  (voxel_grid_covariance_omp_impl.hpp:355,359; gicp_omp_impl.hpp:450)
//   JacobiSVD<6x6>::solve              (ndt_omp_impl.hpp:127-129)
//   JacobiSVD<3x3, ComputeFullU>       (gicp_omp_impl.hpp:110)
//   Matrix3f::eulerAngles(0,1,2)       (ndt_omp_impl.hpp:109)
// Restated from the published algorithms; any backward-stable routine agrees with Eigen to ~1e-15,
// far inside the 1e-3 m / 1e-3 rad pose tolerance.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace oracle {

// ---- 3x3 helpers, row-major double[9] ----
inline void mat3_mul(const double* a, const double* b, double* c) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
      t[i * 3 + j] = s;
    }
  std::memcpy(c, t, sizeof(t));
}

inline void mat3_transpose(const double* a, double* c) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = a[j * 3 + i];
  std::memcpy(c, t, sizeof(t));
}

// General 3x3 inverse by cofactors / determinant (what Eigen does for fixed-size 3x3).
inline void mat3_inverse(const double* m, double* out) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c01 = m[5] * m[6] - m[3] * m[8];
  double c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  double t[9];
  t[0] = c00 * id;
  t[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  t[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  t[3] = c01 * id;
  t[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  t[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  t[6] = c02 * id;
  t[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  t[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  std::memcpy(out, t, sizeof(t));
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi). evals ascending, evecs columns (row-major
// storage: evecs[r*3+c] is component r of eigenvector c) — same convention as
// Eigen::SelfAdjointEigenSolver (eigenvalues sorted in increasing order).
inline void sym_eigen3(const double* a_in, double* evals, double* evecs) {
  double a[9];
  std::memcpy(a, a_in, sizeof(a));
  // symmetrise from the lower triangle (Eigen reads the lower triangle only)
  a[1] = a[3];
  a[2] = a[6];
  a[5] = a[7];
  double v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double apq = a[p * 3 + q];
        if (apq == 0.0) continue;
        double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A * G
          double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = c * akp - s * akq;
          a[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- G^T * A
          double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = c * apk - s * aqk;
          a[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
          v[k * 3 + p] = c * vkp - s * vkq;
          v[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  double d[3] = {a[0], a[4], a[8]};
  std::sort(order, order + 3, [&](int x, int y) { return d[x] < d[y]; });
  for (int c = 0; c < 3; c++) {
    evals[c] = d[order[c]];
    for (int r = 0; r < 3; r++) evecs[r * 3 + c] = v[r * 3 + order[c]];
  }
}

// One-sided (Hestenes) Jacobi SVD of an N x N matrix: A = U diag(s) V^T. Row-major.
template <int N>
struct JacobiSVD {
  double U[N * N], V[N * N], S[N];
  explicit JacobiSVD(const double* A) {
    double W[N * N];
    std::memcpy(W, A, sizeof(W));
    for (int i = 0; i < N * N; i++) V[i] = 0;
    for (int i = 0; i < N; i++) V[i * N + i] = 1;
    const double eps = std::numeric_limits<double>::epsilon();
    for (int sweep = 0; sweep < 80; sweep++) {
      bool rotated = false;
      for (int p = 0; p < N - 1; p++)
        for (int q = p + 1; q < N; q++) {
          double alpha = 0, beta = 0, gamma = 0;
          for (int k = 0; k < N; k++) {
            alpha += W[k * N + p] * W[k * N + p];
            beta += W[k * N + q] * W[k * N + q];
            gamma += W[k * N + p] * W[k * N + q];
          }
          if (gamma == 0.0 || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
          rotated = true;
          double zeta = (beta - alpha) / (2.0 * gamma);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
          for (int k = 0; k < N; k++) {
            double wp = W[k * N + p], wq = W[k * N + q];
            W[k * N + p] = c * wp - s * wq;
            W[k * N + q] = s * wp + c * wq;
            double vp = V[k * N + p], vq = V[k * N + q];
            V[k * N + p] = c * vp - s * vq;
            V[k * N + q] = s * vp + c * vq;
          }
        }
      if (!rotated) break;
    }
    // singular values = column norms; sort descending (Eigen convention)
    int order[N];
    double nrm[N];
    for (int j = 0; j < N; j++) {
      double s = 0;
      for (int k = 0; k < N; k++) s += W[k * N + j] * W[k * N + j];
      nrm[j] = std::sqrt(s);
      order[j] = j;
    }
    std::stable_sort(order, order + N, [&](int x, int y) { return nrm[x] > nrm[y]; });
    double Vs[N * N];
    for (int j = 0; j < N; j++) {
      int o = order[j];
      S[j] = nrm[o];
      for (int k = 0; k < N; k++) {
        U[k * N + j] = nrm[o] > 0 ? W[k * N + o] / nrm[o] : 0.0;
        Vs[k * N + j] = V[k * N + o];
      }
    }
    std::memcpy(V, Vs, sizeof(Vs));
    // complete U to an orthonormal basis for zero singular values (ComputeFullU): Gram-Schmidt on e_i
    for (int j = 0; j < N; j++) {
      if (S[j] > 0) continue;
      for (int e = 0; e < N; e++) {
        double cand[N];
        for (int k = 0; k < N; k++) cand[k] = (k == e) ? 1.0 : 0.0;
        for (int jj = 0; jj < N; jj++) {
          if (jj == j) continue;
          double cn = 0;
          for (int k = 0; k < N; k++) cn += U[k * N + jj] * U[k * N + jj];
          if (cn == 0) continue;
          double d = 0;
          for (int k = 0; k < N; k++) d += cand[k] * U[k * N + jj];
          for (int k = 0; k < N; k++) cand[k] -= d * U[k * N + jj];
        }
        double n2 = 0;
        for (int k = 0; k < N; k++) n2 += cand[k] * cand[k];
        if (n2 > 1e-8) {
          double inv = 1.0 / std::sqrt(n2);
          for (int k = 0; k < N; k++) U[k * N + j] = cand[k] * inv;
          break;
        }
      }
    }
  }
  // Minimum-norm least-squares solve; singular values <= max(N)*eps*s_max treated as zero
  // (Eigen JacobiSVD::solve default threshold).
  void solve(const double* b, double* x) const {
    const double thr = S[0] * double(N) * std::numeric_limits<double>::epsilon();
    for (int i = 0; i < N; i++) x[i] = 0;
    for (int j = 0; j < N; j++) {
      if (!(S[j] > thr)) continue;
      double d = 0;
      for (int k = 0; k < N; k++) d += U[k * N + j] * b[k];
      d /= S[j];
      for (int k = 0; k < N; k++) x[k] += d * V[k * N + j];
    }
  }
};

// Eigen 3.4 MatrixBase<Matrix3f>::eulerAngles(0,1,2) (Graphics Gems IV routine), float arithmetic.
// m is row-major float[9]. Result first angle in [0, pi].
inline void euler_angles_012(const float* m, float* out) {
  const float kPi = 3.14159265358979323846f;
  auto M = [&](int r, int c) { return m[r * 3 + c]; };
  float r0 = std::atan2(M(1, 2), M(2, 2));
  float c2 = std::sqrt(M(0, 0) * M(0, 0) + M(0, 1) * M(0, 1));
  float r1;
  if (r0 > 0.0f) {
    r0 -= kPi;
    r1 = std::atan2(-M(0, 2), -c2);
  } else {
    r1 = std::atan2(-M(0, 2), c2);
  }
  float s1 = std::sin(r0), c1 = std::cos(r0);
  float r2 = std::atan2(s1 * M(2, 0) - c1 * M(1, 0), c1 * M(1, 1) - s1 * M(2, 1));
  out[0] = -r0;
  out[1] = -r1;
  out[2] = -r2;
}

// Translation(x) * AngleAxis(rx, X) * AngleAxis(ry, Y) * AngleAxis(rz, Z) in float
// (ndt_omp_impl.hpp:146-149, 811-814). Output: row-major 4x4 float.
inline void pose_to_matrix_f(const double* p, float* T) {
  float cx = std::cos((float)p[3]), sx = std::sin((float)p[3]);
  float cy = std::cos((float)p[4]), sy = std::sin((float)p[4]);
  float cz = std::cos((float)p[5]), sz = std::sin((float)p[5]);
  float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float A[9], R[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float s = 0;
      for (int k = 0; k < 3; k++) s += Rx[i * 3 + k] * Ry[k * 3 + j];
      A[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float s = 0;
      for (int k = 0; k < 3; k++) s += A[i * 3 + k] * Rz[k * 3 + j];
      R[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
    T[i * 4 + 3] = (float)p[i];
  }
  T[12] = T[13] = T[14] = 0;
  T[15] = 1;
}

}  // namespace oracle
