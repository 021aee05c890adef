// verify.cuh -- two-view geometric verification stage (K2/K3) hooks used by the pair scheduler.
#pragma once
#include "internal.h"

namespace b2m {
void verify_results_init(b2m_results* res, int64_t n_pairs);
// Allocate the verification workspace (points arena, masks, outputs) and upload the cameras.
int verify_prepare(b2m_ctx* ctx, ImageSet& S, int batch, int64_t arena_cap);
// Device arena of slot `s` the compaction kernel fills with (x1, y1, x2, y2) per raw match.
void* verify_points_arena(b2m_ctx* ctx, int s);
// Enqueue verification of batch slot `s` (pairs [p0, p0+nb)) on ctx->stream, after compaction.
int verify_batch_launch(b2m_ctx* ctx, ImageSet& S, const b2m_tvg_opts* tvg, const b2m_sift_opts* sift, int s,
                        int64_t p0, int nb);
// Enqueue the D2H copies of the verification outputs of slot `s` on ctx->copy_stream.
int verify_batch_download(b2m_ctx* ctx, b2m_results* res, int s, int64_t p0, int nb);
// After the copies completed: move staging -> results (applies the controller's write rule, row P3).
int verify_batch_collect(b2m_ctx* ctx, b2m_results* res, int s, int64_t p0, int nb, int min_num_inliers);
// Guided matching buffers of slot `s` (lazily allocated); marks the slot as guided for download / collect.
struct GuidedSlot {
  int32_t* kind;
  float* model;
  uint2* arena;
  unsigned long long* cursor;
  int64_t* off;
  int32_t* cnt;
  unsigned long long* h_cursor;
};
int verify_guided_slot(b2m_ctx* ctx, int s, GuidedSlot* out);
void verify_release(b2m_ctx* ctx);
}  // namespace b2m
