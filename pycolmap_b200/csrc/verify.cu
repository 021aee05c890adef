// verify.cu -- placeholder until the K2/K3 LO-RANSAC kernels land (next milestone).
#include "verify.cuh"

namespace b2m {
void verify_results_init(b2m_results* res, int64_t n_pairs) {
  res->verified = true;
  res->config.assign(n_pairs, B2M_UNDEFINED);
  res->in_off.assign(n_pairs, 0);
  res->in_cnt.assign(n_pairs, 0);
  res->models.assign(27 * n_pairs, 0.0);
}
int verify_batch_launch(b2m_ctx* ctx, ImageSet&, const b2m_tvg_opts*, const b2m_sift_opts*, int, int64_t, int) {
  ctx->err = "[verify.cu] two-view verification kernels not built yet";
  return B2M_ESTATE;
}
int verify_batch_download(b2m_ctx*, b2m_results*, int, int64_t, int) { return B2M_OK; }
int verify_batch_collect(b2m_ctx*, b2m_results*, int, int64_t, int) { return B2M_OK; }
void verify_release(b2m_ctx*) {}
}  // namespace b2m

extern "C" {
int b2m_estimate_two_view_geometry(b2m_ctx* ctx, const b2m_camera*, const double*, int64_t, const b2m_camera*,
                                   const double*, int64_t, const uint32_t*, int64_t, const b2m_tvg_opts*,
                                   b2m_tvg_result*, uint32_t*) {
  if (ctx) ctx->err = "[verify.cu] not built yet";
  return B2M_ESTATE;
}
int b2m_ransac_model(b2m_ctx* ctx, int32_t, const double*, const double*, int64_t, const b2m_ransac_opts*, double*,
                     uint8_t*, int64_t*, int32_t*) {
  if (ctx) ctx->err = "[verify.cu] not built yet";
  return B2M_ESTATE;
}
int b2m_squared_sampson_error(b2m_ctx* ctx, const double*, const double*, int64_t, const double*, double*) {
  if (ctx) ctx->err = "[verify.cu] not built yet";
  return B2M_ESTATE;
}
}
