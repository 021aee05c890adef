// Host build of pycolmap_b200/csrc/pose.h for CPU unit tests (test infrastructure).  ph_select_pose is the
// serial form of the control flow b2m_pose_kernel runs with one CTA per pair.
#include <algorithm>
#include <vector>

#include "../../pycolmap_b200/csrc/pose.h"
using namespace b2m::pose;
extern "C" {
void ph_svd3(const double* A, double* U, double* S, double* V) { svd3(A, U, S, V); }
void ph_decompose_E(const double* E, double* R1, double* R2, double* t) { decompose_E(E, R1, R2, t); }
int ph_decompose_H(const double* H, const double* K1, const double* K2, double* R, double* t, double* n) {
  return decompose_H(H, K1, K2, R, t, n);
}
int ph_triangulate(const double* R, const double* t, double x1, double y1, double x2, double y2, double* X) {
  return triangulate(R, t, x1, y1, x2, y2, X) ? 1 : 0;
}
void ph_quat(const double* R, double* q) { rotation_to_quat(R, q); }
// candidates (R [n_cand x 9], t [n_cand x 3]) + normalised inlier points -> chosen candidate, number of points in
// front of both cameras, median triangulation angle.  Later candidates win ties, the first one is always taken.
int ph_select_pose(const double* R, const double* t, int n_cand, const double* x1, const double* x2, int n, int* n_front,
                   double* tri_angle) {
  int best = 0, best_cnt = -1;
  for (int c = 0; c < n_cand; ++c) {
    const double md = cheirality_max_depth(R + 9 * c, t + 3 * c);
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
      double X[3];
      if (triangulate(R + 9 * c, t + 3 * c, x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1], X) &&
          in_front_of_both(R + 9 * c, t + 3 * c, X, md))
        ++cnt;
    }
    if (c == 0 || cnt >= best_cnt) {
      best = c;
      best_cnt = cnt;
    }
  }
  const double* Rb = R + 9 * best;
  const double* tb = t + 3 * best;
  const double c2[3] = {-(Rb[0] * tb[0] + Rb[3] * tb[1] + Rb[6] * tb[2]), -(Rb[1] * tb[0] + Rb[4] * tb[1] + Rb[7] * tb[2]),
                        -(Rb[2] * tb[0] + Rb[5] * tb[1] + Rb[8] * tb[2])};
  const double md = cheirality_max_depth(Rb, tb);
  std::vector<double> ang;
  for (int i = 0; i < n; ++i) {
    double X[3];
    if (triangulate(Rb, tb, x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1], X) && in_front_of_both(Rb, tb, X, md))
      ang.push_back(triangulation_angle(c2, X));
  }
  std::sort(ang.begin(), ang.end());
  *n_front = static_cast<int>(ang.size());
  const size_t m = ang.size() / 2;
  *tri_angle = ang.empty() ? 0.0 : (ang.size() % 2 == 0 ? 0.5 * (ang[m - 1] + ang[m]) : ang[m]);
  return best;
}
}
