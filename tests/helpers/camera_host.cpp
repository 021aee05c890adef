// Host build of pycolmap_b200/csrc/camera_models.h for CPU unit tests (test infrastructure: the same
// header the CUDA kernels compile is checked against the numpy oracle without a GPU).
#include "../../pycolmap_b200/csrc/camera_models.h"
using namespace b2m::cam;
extern "C" {
int ch_num_params(int model) { return num_params(model); }
double ch_mean_focal_length(int model, const double* p) { return mean_focal_length(model, p); }
void ch_cam_from_img(int model, const double* p, const double* pts, int n, double* out) {
  for (int i = 0; i < n; ++i) cam_from_img(model, p, pts[2 * i], pts[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}
void ch_img_from_cam(int model, const double* p, const double* uv, int n, double* out) {
  for (int i = 0; i < n; ++i) img_from_cam(model, p, uv[2 * i], uv[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}
}
