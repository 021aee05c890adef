// Host build of pycolmap_b200/csrc/geom.h for CPU unit tests (test infrastructure: lets pytest
// check the solver math against numpy without a GPU).  Not part of the product library.
#include "../../pycolmap_b200/csrc/geom.h"
using namespace b2m::geom;
extern "C" {
int gh_poly_roots(const double* c, int deg, double* roots) { return poly_real_roots(c, deg, roots); }
void gh_jacobi9(const double* A_in, double* V, double* w) {
  double A[81];
  for (int i = 0; i < 81; ++i) A[i] = A_in[i];
  jacobi_eig_sym<9>(A, V, w);
}
int gh_estimate_E(const double* x1, const double* y1, const double* x2, const double* y2, int n, double* models) {
  return estimate_E(x1, y1, x2, y2, n, models);
}
int gh_estimate_F7(const double* x1, const double* y1, const double* x2, const double* y2, double* models) {
  return estimate_F7(x1, y1, x2, y2, models);
}
int gh_estimate_F8(const double* x1, const double* y1, const double* x2, const double* y2, int n, double* model) {
  return estimate_F8(x1, y1, x2, y2, n, model);
}
int gh_estimate_H(const double* x1, const double* y1, const double* x2, const double* y2, int n, double* model) {
  return estimate_H(x1, y1, x2, y2, n, model);
}
int gh_five_point_from_nullspace(const double* N, double* models) { return five_point_from_nullspace(N, models); }
int gh_nullspace5(const double* x1, const double* y1, const double* x2, const double* y2, double* N) {
  double A[45];
  for (int i = 0; i < 5; ++i) epipolar_row(x1[i], y1[i], x2[i], y2[i], A + 9 * i);
  return nullspace_gauss<5>(A, N) ? 1 : 0;
}
int gh_minimal_E5(const double* x1, const double* y1, const double* x2, const double* y2, double* m) { return minimal_E5(x1, y1, x2, y2, m); }
int gh_minimal_F7(const double* x1, const double* y1, const double* x2, const double* y2, double* m) { return minimal_F7(x1, y1, x2, y2, m); }
int gh_minimal_H4(const double* x1, const double* y1, const double* x2, const double* y2, double* m) { return minimal_H4(x1, y1, x2, y2, m); }
int gh_minimal_H4_closed(const double* x1, const double* y1, const double* x2, const double* y2, double* m) { return minimal_H4_closed(x1, y1, x2, y2, m); }
// S: symmetric 9x9 (row-major) -> K smallest eigenvectors by inverse subspace iteration
void gh_invit(const double* S, int K, double* out) {
  double S45[45];
  int k = 0;
  for (int i = 0; i < 9; ++i)
    for (int j = i; j < 9; ++j) S45[k++] = S[i * 9 + j];
  if (K == 1) smallest_eigvecs_invit<1>(S45, out);
  if (K == 4) smallest_eigvecs_invit<4>(S45, out);
}
double gh_sampson(const double* E, double x1, double y1, double x2, double y2) { return sampson_sq(E, x1, y1, x2, y2); }
double gh_homography(const double* H, double x1, double y1, double x2, double y2) { return homography_sq(H, x1, y1, x2, y2); }
double gh_num_trials(double ni, double ns, double conf, double mult, int k) { return compute_num_trials(ni, ns, conf, mult, k); }
}
