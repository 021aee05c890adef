// TEST INFRASTRUCTURE ONLY -- a CPU stand-in for libb200match.so, built into tests/helpers/mock/ and put in
// front of the real library with LD_LIBRARY_PATH inside a test subprocess, so that the C++ host layer
// (pycolmap_b200/host: database, pair lists, uploads, gpu_index slicing, write order, resume, verify_matches)
// can be exercised end to end without a GPU.  It is never shipped, never linked by the product and never
// measured.  Matching = the oracle's exact brute-force matcher (oracle/liboracle.so); "verification" is a
// deterministic placeholder (every second match is an inlier, models derived from the pair indices) that
// only has to be recognisable in the database -- the real verifier is tested on the GPU.
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200match.h"

extern "C" int orc_fast_match_pair(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float max_ratio,
                                   float max_distance, int cross_check, uint32_t* out_matches);

// In-process stand-in for an NCCL communicator: the ranks (host threads) deposit their shards in a shared table
// and meet at a barrier; what comes out must be the whole image set on every "GPU".
struct MockComm {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<std::vector<uint8_t>> desc;
  std::vector<std::vector<float>> kpts;
  void barrier(std::unique_lock<std::mutex>& lk) {
    const long gen = generation;
    if (++arrived == n) {
      arrived = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

struct b2m_ctx {
  int device = 0;
  std::shared_ptr<MockComm> comm;
  int comm_rank = 0;
  uint64_t sharded_uploads = 0, sharded_bytes = 0;
  std::string err;
  std::vector<std::vector<uint8_t>> desc;
  std::vector<std::vector<float>> kpts;
  std::vector<b2m_camera> cams;
  volatile int stop = 0;
  uint64_t launches = 0;
};

struct b2m_results {
  std::vector<int32_t> pairs;
  std::vector<std::vector<uint32_t>> matches, inliers;
  std::vector<int32_t> config;
  std::vector<double> models;  // 27 per pair
  bool verified = false;
};

static void fake_geometry(int a, int b, int n_matches, const uint32_t* m, int min_inl, int* config,
                          std::vector<uint32_t>* inl, double* models27) {
  *config = B2M_UNDEFINED;
  inl->clear();
  memset(models27, 0, sizeof(double) * 27);
  if (n_matches < min_inl) return;
  for (int i = 0; i < n_matches; i += 2) {
    inl->push_back(m[2 * i]);
    inl->push_back(m[2 * i + 1]);
  }
  if (static_cast<int>(inl->size() / 2) < min_inl) {
    inl->clear();
    return;
  }
  *config = B2M_CALIBRATED;
  for (int k = 0; k < 9; ++k) {
    models27[k] = 100.0 * a + b + 0.125 * k;             // E: depends on the ORDER of the pair
    models27[9 + k] = -(100.0 * a + b) - 0.25 * k;       // F
    models27[18 + k] = (k % 4 == 0 ? 2.0 + a : 0.0) + 0.0625 * k * (b + 1);  // H: invertible
  }
}

extern "C" {

int b2m_abi_version(void) { return B2M_ABI_VERSION; }
int b2m_device_count(void) {   // MOCK_B2M_DEVICES "GPUs" (default 1), at most 8
  const char* e = getenv("MOCK_B2M_DEVICES");
  const int n = e ? atoi(e) : 1;
  return n < 0 ? 0 : (n > 8 ? 8 : n);
}

int b2m_create(const b2m_device_cfg* cfg, b2m_ctx** out) {
  if (!out) return B2M_EINVAL;
  b2m_ctx* c = new b2m_ctx();
  c->device = cfg ? cfg->device : 0;
  if (c->device < 0 || c->device > 7) {  // the mock box has 8 "GPUs"
    delete c;
    return B2M_ENODEV;
  }
  *out = c;
  return B2M_OK;
}
void b2m_destroy(b2m_ctx* ctx) { delete ctx; }
const char* b2m_last_error(const b2m_ctx* ctx) { return ctx ? ctx->err.c_str() : "mock: create failed"; }
int b2m_request_stop(b2m_ctx* ctx) {
  if (ctx) ctx->stop = 1;
  return B2M_OK;
}

void b2m_sift_opts_default(b2m_sift_opts* o) {
  memset(o, 0, sizeof(*o));
  o->struct_size = sizeof(*o);
  o->max_ratio = 0.8f; o->max_distance = 0.7f; o->cross_check = 1; o->max_num_matches = 32768;
}
void b2m_ransac_opts_default(b2m_ransac_opts* o) {
  memset(o, 0, sizeof(*o));
  o->struct_size = sizeof(*o);
  o->min_num_trials = 100; o->max_num_trials = 10000; o->max_error = 4.0; o->min_inlier_ratio = 0.25;
  o->confidence = 0.999; o->dyn_num_trials_multiplier = 3.0;
}
void b2m_tvg_opts_default(b2m_tvg_opts* o) {
  memset(o, 0, sizeof(*o));
  o->struct_size = sizeof(*o);
  o->min_num_inliers = 15; o->min_E_F_inlier_ratio = 0.95; o->max_H_inlier_ratio = 0.8;
  o->watermark_min_inlier_ratio = 0.7; o->watermark_border_size = 0.1; o->detect_watermark = 1;
  o->multiple_ignore_watermark = 1;
  b2m_ransac_opts_default(&o->ransac);
}

int b2m_set_images(b2m_ctx* ctx, int32_t n, const int32_t* n_feat, const uint8_t* const* desc, const float* const* kpts,
                   const b2m_camera* cams) {
  ctx->desc.assign(n, {});
  ctx->kpts.assign(n, {});
  ctx->cams.clear();
  for (int i = 0; i < n; ++i) {
    ctx->desc[i].assign(desc[i], desc[i] + static_cast<size_t>(n_feat[i]) * 128);
    if (kpts) ctx->kpts[i].assign(kpts[i], kpts[i] + static_cast<size_t>(n_feat[i]) * 2);
  }
  if (cams) {
    for (int i = 0; i < n; ++i)
      if (cams[i].struct_size != sizeof(b2m_camera)) {
        ctx->err = "mock: b2m_camera.struct_size";
        return B2M_EINVAL;
      }
    ctx->cams.assign(cams, cams + n);
  }
  return B2M_OK;
}
int b2m_comm_get_unique_id(b2m_comm_id*) { return B2M_ENODEV; }   // the mock has no cross-process transport
int b2m_comm_init_rank(b2m_ctx* ctx, int32_t, int32_t, const b2m_comm_id*) {
  ctx->err = "mock: no NCCL";
  return B2M_ENODEV;
}
int b2m_comm_init_local(b2m_ctx* const* ctxs, int32_t n) {
  if (getenv("MOCK_B2M_NO_NCCL")) return B2M_ENODEV;           // the host must then upload the whole set everywhere
  auto comm = std::make_shared<MockComm>();
  comm->n = n;
  for (int i = 0; i < n; ++i) {
    ctxs[i]->comm = comm;
    ctxs[i]->comm_rank = i;
  }
  return B2M_OK;
}
int b2m_comm_destroy(b2m_ctx* ctx) {
  ctx->comm.reset();
  return B2M_OK;
}
void b2m_comm_image_range(int32_t n_images, int32_t n_ranks, int32_t rank, int32_t* first, int32_t* count) {
  if (n_ranks < 1) n_ranks = 1;
  const int64_t per = (static_cast<int64_t>(n_images) + n_ranks - 1) / n_ranks;
  const int64_t lo = per * rank < n_images ? per * rank : n_images, hi = per * (rank + 1) < n_images ? per * (rank + 1) : n_images;
  *first = static_cast<int32_t>(lo);
  *count = static_cast<int32_t>(hi - lo);
}
int b2m_set_images_sharded(b2m_ctx* ctx, int32_t n, const int32_t* n_feat, const b2m_camera* cams, const b2m_image_shard* mine) {
  if (!mine || mine->struct_size != sizeof(b2m_image_shard) || mine->location != B2M_LOC_HOST) {
    ctx->err = "mock: shard";
    return B2M_EINVAL;
  }
  const int n_ranks = ctx->comm ? ctx->comm->n : 1;
  int32_t first = 0, count = 0;
  b2m_comm_image_range(n, n_ranks, ctx->comm_rank, &first, &count);
  if (first != mine->first_image || count != mine->n_local) {
    ctx->err = "mock: the shard is not this rank's image range";
    return B2M_EINVAL;
  }
  std::shared_ptr<MockComm> local;
  MockComm* c = ctx->comm.get();
  if (!c) {
    local = std::make_shared<MockComm>();
    local->n = 1;
    c = local.get();
  }
  std::unique_lock<std::mutex> lk(c->m);
  if (static_cast<int>(c->desc.size()) != n) {
    c->desc.assign(n, {});
    c->kpts.assign(n, {});
  }
  const uint8_t* d = static_cast<const uint8_t*>(mine->desc_packed);
  const float* k = static_cast<const float*>(mine->kpts_packed);
  size_t row = 0;
  for (int i = first; i < first + count; ++i) {
    c->desc[i].assign(d + row * 128, d + (row + n_feat[i]) * 128);
    if (mine->has_keypoints) c->kpts[i].assign(k + row * 2, k + (row + n_feat[i]) * 2);
    row += n_feat[i];
    ctx->sharded_bytes += static_cast<uint64_t>(n_feat[i]) * 128;
  }
  c->barrier(lk);                 // the "all-gather"
  ctx->desc = c->desc;
  ctx->kpts = c->kpts;
  for (int i = 0; i < n; ++i)
    if (ctx->desc[i].size() != static_cast<size_t>(n_feat[i]) * 128) {
      ctx->err = "mock: an image is missing after the gather";
      return B2M_ESTATE;
    }
  c->barrier(lk);                 // nobody overwrites the table before everybody has copied it
  ctx->cams.clear();
  if (cams) ctx->cams.assign(cams, cams + n);
  ctx->sharded_uploads += 1;
  return B2M_OK;
}
int b2m_set_images_device(b2m_ctx* ctx, int32_t, const int32_t*, const void*, const void*, const b2m_camera*) {
  ctx->err = "mock: no device memory";
  return B2M_ESTATE;
}

int b2m_match_pairs(b2m_ctx* ctx, const int32_t* pairs, int64_t n_pairs, const b2m_sift_opts* sift, const b2m_tvg_opts* tvg,
                    b2m_results** out) {
  if (ctx->stop) {
    ctx->stop = 0;
    ctx->err = "mock: stopped";
    return B2M_ESTOPPED;
  }
  b2m_results* r = new b2m_results();
  r->pairs.assign(pairs, pairs + 2 * n_pairs);
  r->matches.resize(n_pairs);
  r->inliers.resize(n_pairs);
  r->config.assign(n_pairs, B2M_UNDEFINED);
  r->models.assign(27 * n_pairs, 0.0);
  r->verified = tvg != nullptr;
  for (int64_t k = 0; k < n_pairs; ++k) {
    const int a = pairs[2 * k], b = pairs[2 * k + 1];
    if (a < 0 || b < 0 || a >= static_cast<int>(ctx->desc.size()) || b >= static_cast<int>(ctx->desc.size())) {
      delete r;
      ctx->err = "mock: pair index out of range";
      return B2M_EINVAL;
    }
    const int n1 = static_cast<int>(ctx->desc[a].size() / 128), n2 = static_cast<int>(ctx->desc[b].size() / 128);
    std::vector<uint32_t> m(static_cast<size_t>(n1 > 0 ? n1 : 1) * 2);
    const int n = orc_fast_match_pair(ctx->desc[a].data(), n1, ctx->desc[b].data(), n2, sift->max_ratio, sift->max_distance,
                                      sift->cross_check, m.data());
    m.resize(static_cast<size_t>(n) * 2);
    if (tvg) {
      int cfg;
      fake_geometry(a, b, n, m.data(), tvg->min_num_inliers, &cfg, &r->inliers[k], r->models.data() + 27 * k);
      r->config[k] = cfg;
      if (n < tvg->min_num_inliers) m.clear();  // the library's write rule (row P3)
    }
    r->matches[k] = std::move(m);
    ++ctx->launches;
  }
  *out = r;
  return B2M_OK;
}
int b2m_match_verify(b2m_ctx* ctx, const int32_t* pairs, int64_t n_pairs, const b2m_sift_opts* sift, const b2m_tvg_opts* tvg,
                     b2m_results** out) {
  return b2m_match_pairs(ctx, pairs, n_pairs, sift, tvg, out);
}
int b2m_match_pair(b2m_ctx*, const uint8_t* d1, int32_t n1, const uint8_t* d2, int32_t n2, const b2m_sift_opts* o,
                   uint32_t* out, int64_t, int64_t* out_n) {
  *out_n = orc_fast_match_pair(d1, n1, d2, n2, o->max_ratio, o->max_distance, o->cross_check, out);
  return B2M_OK;
}

int64_t b2m_results_num_pairs(const b2m_results* r) { return r ? static_cast<int64_t>(r->matches.size()) : 0; }
int64_t b2m_results_total_matches(const b2m_results* r) {
  int64_t t = 0;
  if (r) for (const auto& m : r->matches) t += static_cast<int64_t>(m.size() / 2);
  return t;
}
int64_t b2m_results_num_verified(const b2m_results* r) {
  int64_t t = 0;
  if (r) for (int c : r->config) t += c != B2M_UNDEFINED;
  return t;
}
int b2m_results_get(const b2m_results* r, int64_t k, b2m_pair_view* out) {
  if (!r || k < 0 || k >= static_cast<int64_t>(r->matches.size())) return B2M_EINVAL;
  memset(out, 0, sizeof(*out));
  out->struct_size = sizeof(*out);
  out->qvec[0] = 1.0;
  out->image1 = r->pairs[2 * k];
  out->image2 = r->pairs[2 * k + 1];
  out->config = r->config[k];
  out->n_matches = static_cast<int64_t>(r->matches[k].size() / 2);
  out->matches = r->matches[k].empty() ? nullptr : r->matches[k].data();
  out->n_inliers = static_cast<int64_t>(r->inliers[k].size() / 2);
  out->inlier_matches = r->inliers[k].empty() ? nullptr : r->inliers[k].data();
  memcpy(out->E, r->models.data() + 27 * k, 72);
  memcpy(out->F, r->models.data() + 27 * k + 9, 72);
  memcpy(out->H, r->models.data() + 27 * k + 18, 72);
  return B2M_OK;
}
void b2m_results_free(b2m_results* r) { delete r; }

int b2m_estimate_two_view_geometry(b2m_ctx*, const b2m_camera*, const double*, int64_t n1, const b2m_camera*, const double*,
                                   int64_t, const uint32_t* matches, int64_t m, const b2m_tvg_opts* opts, b2m_tvg_result* out,
                                   uint32_t* inlier_matches) {
  std::vector<uint32_t> ident;
  if (!matches) {
    m = n1;
    for (int64_t i = 0; i < m; ++i) { ident.push_back(static_cast<uint32_t>(i)); ident.push_back(static_cast<uint32_t>(i)); }
    matches = ident.data();
  }
  std::vector<uint32_t> inl;
  double models[27];
  int cfg;
  fake_geometry(7, 9, static_cast<int>(m), matches, opts->min_num_inliers, &cfg, &inl, models);
  memset(out, 0, sizeof(*out));
  out->struct_size = sizeof(*out);
  out->qvec[0] = 1.0;
  out->config = cfg == B2M_UNDEFINED ? B2M_DEGENERATE : cfg;
  out->n_inliers = static_cast<int64_t>(inl.size() / 2);
  memcpy(out->E, models, 72); memcpy(out->F, models + 9, 72); memcpy(out->H, models + 18, 72);
  if (!inl.empty()) memcpy(inlier_matches, inl.data(), inl.size() * 4);
  return B2M_OK;
}
int b2m_estimate_two_view_geometry_batch(b2m_ctx* ctx, const b2m_tvg_problem* q, int64_t n, const b2m_tvg_opts* opts,
                                         b2m_tvg_result* out, uint32_t* const* inl) {
  for (int64_t k = 0; k < n; ++k)
    if (int rc = b2m_estimate_two_view_geometry(ctx, &q[k].cam1, q[k].points1, q[k].n1, &q[k].cam2, q[k].points2, q[k].n2,
                                                q[k].matches, q[k].m, opts, &out[k], inl ? inl[k] : nullptr))
      return rc;
  return B2M_OK;
}
int b2m_estimate_two_view_geometry_pose(b2m_ctx*, const b2m_camera*, const double*, int64_t, const b2m_camera*, const double*,
                                        int64_t, const uint32_t*, int64_t, b2m_tvg_result* g) {
  g->qvec[0] = 1.0;
  g->pose_valid = 0;
  return B2M_OK;
}
int b2m_ransac_model(b2m_ctx*, int32_t, const double*, const double*, int64_t, const b2m_ransac_opts*, double*, uint8_t*,
                     int64_t* n, int32_t* ok) {
  *n = 0;
  *ok = 0;
  return B2M_OK;
}
int b2m_debug_five_point(b2m_ctx* ctx, const double*, int64_t, double*, int32_t*) {
  ctx->err = "mock: no solver";
  return B2M_ESTATE;
}
int b2m_cam_from_img(b2m_ctx*, const b2m_camera*, const double* p, int64_t n, double* out) {
  memcpy(out, p, sizeof(double) * 2 * n);
  return B2M_OK;
}
int b2m_squared_sampson_error(b2m_ctx*, const double*, const double*, int64_t m, const double*, double* out) {
  for (int64_t i = 0; i < m; ++i) out[i] = 0.0;
  return B2M_OK;
}
int b2m_get_stats(b2m_ctx* ctx, b2m_stats* out) {
  memset(out, 0, sizeof(*out));
  out->struct_size = sizeof(*out);
  out->kernel_launches = ctx->launches;
  out->last_allgather_bytes = ctx->sharded_bytes;   // bytes this "GPU" uploaded itself (the host's shard)
  out->comm_size = ctx->comm ? ctx->comm->n : 1;
  out->comm_rank = ctx->comm_rank;
  return B2M_OK;
}
int b2m_reset_stats(b2m_ctx* ctx) {
  ctx->launches = 0;
  return B2M_OK;
}
}
