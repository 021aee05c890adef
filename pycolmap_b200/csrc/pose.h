// pose.h -- relative pose from a verified two-view geometry (TwoViewGeometryOptions.compute_relative_pose):
// essential-matrix decomposition + cheirality, homography decomposition, triangulation, triangulation
// angles.  Header-only, fp64, host + device (unit-tested on the CPU through tests/helpers/pose_host.cpp).
//
// Restates (from the published algorithms; COLMAP 3.9.1 is not on disk -- parity unpinned):
//   U:geometry/essential_matrix.cc   DecomposeEssentialMatrix, PoseFromEssentialMatrix
//   U:geometry/homography_matrix.cc  DecomposeHomographyMatrix (Malis & Vargas, "Deeper understanding of the
//                                    homography decomposition for vision-based control", analytical method),
//                                    PoseFromHomographyMatrix
//   U:geometry/pose.cc               CheckCheirality;  U:geometry/triangulation.cc TriangulatePoint,
//                                    CalculateTriangulationAngle;  U:scene/projection.cc CalculateDepth
// reached from R:estimators/two_view_geometry.h:153-158 (estimate_two_view_geometry_pose) and from
// EstimateTwoViewGeometry when options.compute_relative_pose is set (R:estimators/two_view_geometry.h:57).
// Matrices are row-major; a pose is x_cam2 = R * x_cam1 + t.
#pragma once
#include "geom.h"

namespace b2m {
namespace pose {

using geom::jacobi_eig_sym;

B2M_HD inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
B2M_HD inline void mat3_transpose(const double* A, double* T) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}
B2M_HD inline double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
B2M_HD inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
B2M_HD inline double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Thin SVD of a 3x3 matrix through the eigen-decomposition of A^T A: A = U diag(S) V^T with S descending,
// U and V proper rotations' columns up to a common sign (det(U) = det(V) = +1 is enforced by flipping the
// third columns, which leaves A = U S V^T intact when S[2] == 0 and flips S[2]'s sign otherwise -- callers
// here only need rank-2 inputs or the singular VALUES).
B2M_HD inline void svd3(const double* A, double* U, double* S, double* V) {
  double AtA[9], Vv[9], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) AtA[i * 3 + j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
  jacobi_eig_sym<3>(AtA, Vv, w);
  int o[3] = {0, 1, 2};  // eigenvalues descending
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (w[o[j]] > w[o[i]]) {
        const int t = o[i];
        o[i] = o[j];
        o[j] = t;
      }
  double v[3][3], u[3][3];
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < 3; ++r) v[c][r] = Vv[r * 3 + o[c]];
    S[c] = sqrt(fmax(w[o[c]], 0.0));
  }
  // v3 = v1 x v2 makes V a proper rotation whatever signs the eigen-solver picked
  cross3(v[0], v[1], v[2]);
  const double scale = fmax(S[0], 1e-300);
  for (int c = 0; c < 2; ++c) {
    for (int r = 0; r < 3; ++r) u[c][r] = A[r * 3] * v[c][0] + A[r * 3 + 1] * v[c][1] + A[r * 3 + 2] * v[c][2];
    if (c == 1) {  // orthogonalise against u1 (matters when S[1] is tiny)
      const double d = u[1][0] * u[0][0] + u[1][1] * u[0][1] + u[1][2] * u[0][2];
      for (int r = 0; r < 3; ++r) u[1][r] -= d * u[0][r];
    }
    const double n = norm3(u[c]);
    if (n > 1e-12 * scale && n > 0.0) {
      for (int r = 0; r < 3; ++r) u[c][r] /= n;
    } else if (c == 0) {  // zero matrix
      u[0][0] = 1.0; u[0][1] = 0.0; u[0][2] = 0.0;
    } else {              // rank 1: any unit vector orthogonal to u1
      const double ax[3] = {fabs(u[0][0]) < 0.9 ? 1.0 : 0.0, fabs(u[0][0]) < 0.9 ? 0.0 : 1.0, 0.0};
      cross3(u[0], ax, u[1]);
      const double m = norm3(u[1]);
      for (int r = 0; r < 3; ++r) u[1][r] /= m;
    }
  }
  cross3(u[0], u[1], u[2]);
  {  // sign of the third singular value under the det = +1 convention
    const double Av3[3] = {A[0] * v[2][0] + A[1] * v[2][1] + A[2] * v[2][2], A[3] * v[2][0] + A[4] * v[2][1] + A[5] * v[2][2],
                           A[6] * v[2][0] + A[7] * v[2][1] + A[8] * v[2][2]};
    if (Av3[0] * u[2][0] + Av3[1] * u[2][1] + Av3[2] * u[2][2] < 0.0) S[2] = -S[2];
  }
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      U[r * 3 + c] = u[c][r];
      V[r * 3 + c] = v[c][r];
    }
}

// DecomposeEssentialMatrix: E = U diag(s, s, 0) V^T  ->  R1 = U W V^T, R2 = U W^T V^T, t = u3 (unit).
B2M_HD inline void decompose_E(const double* E, double* R1, double* R2, double* t) {
  double U[9], S[3], V[9], Vt[9], UW[9];
  svd3(E, U, S, V);
  mat3_transpose(V, Vt);
  const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  mat3_mul(U, W, UW);
  mat3_mul(UW, Vt, R1);
  mat3_mul(U, Wt, UW);
  mat3_mul(UW, Vt, R2);
  t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
  const double n = norm3(t);
  if (n > 0.0) {
    t[0] /= n; t[1] /= n; t[2] /= n;
  }
}

// TriangulatePoint for P1 = [I | 0], P2 = [R | t] on normalised image points: null vector of the 4x4 DLT
// system (smallest eigenvector of A^T A).  Returns false when the point is at infinity (w == 0).
B2M_HD inline bool triangulate(const double* R, const double* t, double x1, double y1, double x2, double y2, double* X) {
  double A[4][4];
  // x1 * P1.row(2) - P1.row(0), y1 * P1.row(2) - P1.row(1)
  A[0][0] = -1.0; A[0][1] = 0.0; A[0][2] = x1; A[0][3] = 0.0;
  A[1][0] = 0.0; A[1][1] = -1.0; A[1][2] = y1; A[1][3] = 0.0;
  for (int c = 0; c < 3; ++c) {
    A[2][c] = x2 * R[6 + c] - R[c];
    A[3][c] = y2 * R[6 + c] - R[3 + c];
  }
  A[2][3] = x2 * t[2] - t[0];
  A[3][3] = y2 * t[2] - t[1];
  double AtA[16], V[16], w[4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) AtA[i * 4 + j] = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j] + A[3][i] * A[3][j];
  jacobi_eig_sym<4>(AtA, V, w);
  int k = 0;
  for (int i = 1; i < 4; ++i)
    if (w[i] < w[k]) k = i;
  const double W = V[12 + k];
  if (W == 0.0) return false;
  X[0] = V[k] / W; X[1] = V[4 + k] / W; X[2] = V[8 + k] / W;
  return true;
}

// CheckCheirality's per-point test: depth > eps and < max_depth in both cameras (CalculateDepth = third row of
// [R | t] applied to X, times the norm of the third column of R = 1 for a rotation).
B2M_HD inline bool in_front_of_both(const double* R, const double* t, const double* X, double max_depth) {
  const double kMinDepth = 2.220446049250313e-16;
  const double d1 = X[2];
  if (!(d1 > kMinDepth && d1 < max_depth)) return false;
  const double cn = sqrt(R[2] * R[2] + R[5] * R[5] + R[8] * R[8]);
  const double d2 = (R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2]) * cn;
  return d2 > kMinDepth && d2 < max_depth;
}
// max_depth = 1000 * |R^T t|
B2M_HD inline double cheirality_max_depth(const double* R, const double* t) {
  const double b[3] = {R[0] * t[0] + R[3] * t[1] + R[6] * t[2], R[1] * t[0] + R[4] * t[1] + R[7] * t[2],
                       R[2] * t[0] + R[5] * t[1] + R[8] * t[2]};
  return 1000.0 * norm3(b);
}

// CalculateTriangulationAngle between the rays from the two projection centres (c1 = 0, c2 = -R^T t) to X.
B2M_HD inline double triangulation_angle(const double* c2, const double* X) {
  const double kPi = 3.14159265358979323846;
  const double baseline2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
  const double ray1 = X[0] * X[0] + X[1] * X[1] + X[2] * X[2];
  const double d[3] = {X[0] - c2[0], X[1] - c2[1], X[2] - c2[2]};
  const double ray2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const double denominator = 2.0 * sqrt(ray1 * ray2);
  if (denominator == 0.0) return 0.0;
  const double nominator = ray1 + ray2 - baseline2;
  const double angle = fabs(acos(nominator / denominator));
  return fmin(angle, kPi - angle);
}

// DecomposeHomographyMatrix (analytical method).  H maps image 1 to image 2 pixels; K1, K2 = calibration matrices
// as (fx, fy, cx, cy).  Writes up to 4 candidates (R [9], t [3], n [3] each) and returns their number: 1 for a
// pure rotation (t = n = 0), else 4.
B2M_HD inline int decompose_H(const double* H, const double* K1, const double* K2, double* R_out, double* t_out, double* n_out) {
  // H_normalized = K2^-1 * H * K1
  double Hn[9];
  {
    const double K1m[9] = {K1[0], 0, K1[2], 0, K1[1], K1[3], 0, 0, 1};
    const double K2i[9] = {1.0 / K2[0], 0, -K2[2] / K2[0], 0, 1.0 / K2[1], -K2[3] / K2[1], 0, 0, 1};
    double T[9];
    mat3_mul(K2i, H, T);
    mat3_mul(T, K1m, Hn);
  }
  {  // remove the scale: divide by the middle singular value
    double U[9], S[3], V[9];
    svd3(Hn, U, S, V);
    const double s1 = fabs(S[1]) > 0.0 ? S[1] : 1.0;
    for (int i = 0; i < 9; ++i) Hn[i] /= s1;
  }
  if (det3(Hn) < 0.0)  // always rotations, never reflections
    for (int i = 0; i < 9; ++i) Hn[i] = -Hn[i];
  double S[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S[i * 3 + j] = Hn[i] * Hn[j] + Hn[3 + i] * Hn[3 + j] + Hn[6 + i] * Hn[6 + j] - (i == j ? 1.0 : 0.0);
  double inf_norm = 0.0;  // Eigen's lpNorm<Infinity>() of a matrix: the largest |coefficient|
  for (int i = 0; i < 9; ++i) inf_norm = fmax(inf_norm, fabs(S[i]));
  if (inf_norm < 1e-3) {  // H is a rotation
    for (int i = 0; i < 9; ++i) R_out[i] = Hn[i];
    for (int i = 0; i < 3; ++i) t_out[i] = n_out[i] = 0.0;
    return 1;
  }
  // opposites of the minors of S
  auto minor = [&](int row, int col) {
    const int c0 = col == 0 ? 1 : 0, c1 = col == 2 ? 1 : 2, r0 = row == 0 ? 1 : 0, r1 = row == 2 ? 1 : 2;
    return S[r0 * 3 + c1] * S[r1 * 3 + c0] - S[r0 * 3 + c0] * S[r1 * 3 + c1];
  };
  auto sgn = [](double v) { return v < 0.0 ? -1.0 : 1.0; };
  const double M00 = minor(0, 0), M11 = minor(1, 1), M22 = minor(2, 2);
  const double rtM00 = sqrt(fmax(M00, 0.0)), rtM11 = sqrt(fmax(M11, 0.0)), rtM22 = sqrt(fmax(M22, 0.0));
  const double M01 = minor(0, 1), M12 = minor(1, 2), M02 = minor(0, 2);
  const double e12 = sgn(M12), e02 = sgn(M02), e01 = sgn(M01);
  const double nS00 = fabs(S[0]), nS11 = fabs(S[4]), nS22 = fabs(S[8]);
  int idx = 0;
  if (nS00 < nS11) idx = nS11 < nS22 ? 2 : 1;
  else idx = nS00 < nS22 ? 2 : 0;
  double np1[3], np2[3];
  if (idx == 0) {
    np1[0] = S[0]; np2[0] = S[0];
    np1[1] = S[1] + rtM22; np2[1] = S[1] - rtM22;
    np1[2] = S[2] + e12 * rtM11; np2[2] = S[2] - e12 * rtM11;
  } else if (idx == 1) {
    np1[0] = S[1] + rtM22; np2[0] = S[1] - rtM22;
    np1[1] = S[4]; np2[1] = S[4];
    np1[2] = S[5] - e02 * rtM00; np2[2] = S[5] + e02 * rtM00;
  } else {
    np1[0] = S[2] + e01 * rtM11; np2[0] = S[2] - e01 * rtM11;
    np1[1] = S[5] + rtM00; np2[1] = S[5] - rtM00;
    np1[2] = S[8]; np2[2] = S[8];
  }
  const double traceS = S[0] + S[4] + S[8];
  const double v = 2.0 * sqrt(fmax(1.0 + traceS - M00 - M11 - M22, 0.0));
  const double ESii = sgn(S[idx * 3 + idx]);
  const double r_2 = 2.0 + traceS + v, nt_2 = 2.0 + traceS - v;
  const double r = sqrt(fmax(r_2, 0.0)), n_t = sqrt(fmax(nt_2, 0.0));
  double n1[3], n2[3];
  const double l1 = norm3(np1), l2 = norm3(np2);
  for (int i = 0; i < 3; ++i) {
    n1[i] = np1[i] / l1;
    n2[i] = np2[i] / l2;
  }
  const double half_nt = 0.5 * n_t, esii_t_r = ESii * r;
  double t1s[3], t2s[3];
  for (int i = 0; i < 3; ++i) {
    t1s[i] = half_nt * (esii_t_r * n2[i] - n_t * n1[i]);
    t2s[i] = half_nt * (esii_t_r * n1[i] - n_t * n2[i]);
  }
  // ComputeHomographyRotation: R = H_n * (I - (2 / v) * t_star * n^T)
  auto rotation = [&](const double* ts, const double* n, double* R) {
    double M[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[i * 3 + j] = (i == j ? 1.0 : 0.0) - (2.0 / v) * ts[i] * n[j];
    mat3_mul(Hn, M, R);
  };
  double Ra[9], Rb[9], ta[3], tb[3];
  rotation(t1s, n1, Ra);
  rotation(t2s, n2, Rb);
  for (int i = 0; i < 3; ++i) {
    ta[i] = Ra[i * 3] * t1s[0] + Ra[i * 3 + 1] * t1s[1] + Ra[i * 3 + 2] * t1s[2];
    tb[i] = Rb[i * 3] * t2s[0] + Rb[i * 3 + 1] * t2s[1] + Rb[i * 3 + 2] * t2s[2];
  }
  for (int c = 0; c < 4; ++c) {
    const double* R = c < 2 ? Ra : Rb;
    const double* t = c < 2 ? ta : tb;
    const double* n = c < 2 ? n1 : n2;
    const double ts = (c % 2 == 0) ? 1.0 : -1.0;  // t: {t1, -t1, t2, -t2},  n: {-n1, n1, -n2, n2}
    for (int i = 0; i < 9; ++i) R_out[c * 9 + i] = R[i];
    for (int i = 0; i < 3; ++i) {
      t_out[c * 3 + i] = ts * t[i];
      n_out[c * 3 + i] = -ts * n[i];
    }
  }
  return 4;
}

// Rotation matrix -> unit quaternion (w, x, y, z), w >= 0 branch selection as Eigen's Quaterniond(Matrix3d).
B2M_HD inline void rotation_to_quat(const double* R, double* q) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (R[7] - R[5]) * s;
    q[2] = (R[2] - R[6]) * s;
    q[3] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[1 + i] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (R[k * 3 + j] - R[j * 3 + k]) * s;
    q[1 + j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
    q[1 + k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
  }
}

}  // namespace pose
}  // namespace b2m
