// match_kernel.cuh -- launch interface of the K1 matching kernels (internal to libb200match.so).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace b2m {

struct MatchParams {
  const int32_t* pairs;      // device [n_pairs x 2] image indices of this batch
  const int32_t* img_row0;   // device [n_images] first (padded) row of each image in the descriptor array
  const int32_t* img_nfeat;  // device [n_images] valid rows of each image
  int32_t* mbuf;             // device [n_pairs][2][mstride]: m12 / m21 (column index or -1)
  int32_t mstride;           // >= max padded feature count
  const float* acos_lut;     // device [262145] acosf(min(d * 2^-18, 1)) built with the host libm
  float max_ratio;
  float max_distance;
  // K1 v2 (filter + resolve) only:
  uint2* aux;                // device [n_pairs][2][mstride]: (best, S1) of candidate rows
  int32_t* cand_cnt;         // device [n_pairs][2] number of candidate rows
  int32_t* cand_rows;        // device [n_pairs][2][mstride] candidate row list
  int32_t* cand_sorted;      // device [n_pairs][2][mstride] candidate rows bucketed by winning slot
  // persistent K1: filled in by launch_k1_filter
  int32_t n_items, blocks_per_image, n_dirs;
  // gathered column direction (launch_k1_filter_gather): the "row image" of a work item is the gathered block of the
  // pair (the descriptors of image b that some row of image a matched), the "column image" is image a
  const int32_t* item_list;   // device [n][2] (pair, 256-row block) written by the gather kernel, or nullptr
  const int32_t* n_items_ptr; // device count of item_list entries
  const int32_t* gath_cnt;    // device [n_pairs] gathered rows of each pair
  const uint8_t* gath_desc;   // device [n_pairs x mstride x 128] gathered descriptors (resolve kernel)
};

struct CompactParams {
  const int32_t* pairs;
  const int32_t* img_nfeat;
  const int32_t* mbuf;
  int32_t mstride;
  int32_t cross_check;
  uint2* arena;                 // device match arena of this batch
  unsigned long long* cursor;   // device arena cursor (matches)
  int64_t* pair_off;            // device [n_pairs] offset (in matches) of each pair inside the arena
  int32_t* pair_cnt;            // device [n_pairs]
  // verification input (optional): pixel coordinates of every raw match, same arena offsets
  const int32_t* img_row0;      // device [n_images]
  const float2* kpts;           // device keypoints indexed by padded row, or nullptr
  double4* pts;                 // device arena (x1, y1, x2, y2), or nullptr
  const int32_t* enable;        // optional [n_pairs]: pairs with enable[pair] < 0 produce no output (guided pass)
  const int32_t* colrank;       // optional [n_pairs][mstride]: m21 is indexed by the RANK of a column among the pair's
                                // matched columns (gathered column direction), not by the column itself
  const int32_t* cand_cnt;      // optional [n_pairs][2] candidate counts of the filter epilogue (K1 v2): a pair without a
                                // row-direction candidate has no match, its m12 rows need not be read
};

// Guided matching (K1g): per pair of the batch the geometry chosen by the verifier.
struct GuidedParams {
  const int32_t* kind;          // device [n_pairs]: -1 = not eligible, 0 = F (Sampson), 1 = H (forward transfer)
  const float* model;           // device [n_pairs][9] float32 row-major (as upstream: Eigen::Matrix3f)
  const float2* kpts;           // device keypoints indexed by padded row
  float max_residual;           // max_error^2 in float32
  // gathered column direction (launch_k1_guided_gather), else nullptr: the rows of the launch are the features of
  // image 2 some row of the first launch matched, ranked ascending -- gath_cnt [n_pairs] = rows per pair,
  // gath_cols [n_pairs][mstride] = gathered row -> feature index in image 2
  const int32_t* gath_cnt;
  const int32_t* gath_cols;
  int32_t only_dir;             // -1: blockIdx.y is the direction; 0 / 1: the launch computes this direction only
};

// Rows per A strip / columns per B tile: images are padded (with zero descriptors, which can
// never win a strict `>` comparison against the initial best = 0) to a multiple of this.
constexpr int kRowPad = 256;

cudaError_t launch_k1_match(const CUtensorMap& tmap, const MatchParams& p, int n_pairs, int max_strips, int n_dirs,
                            cudaStream_t stream);
// K1 v2: filter epilogue (slot maxima) + exact dp4a resolution of the candidate rows.  Bit-identical
// results to launch_k1_match; requires a monotone (non-increasing) acos LUT.
cudaError_t launch_k1_filter(const CUtensorMap& tmap, const MatchParams& p,
                             const uint8_t* desc, int n_pairs, int max_strips, int n_dirs, int num_sms,
                             cudaStream_t stream, cudaEvent_t after_filter);
// Cross-check variant of launch_k1_filter that runs the column direction only where it can matter:
//   1. GEMM + filter over all pairs, row direction only (m12 candidates);
//   2. pairs without a single row-direction candidate cannot produce a match whatever m21 holds
//      (FindBestMatchesBruteForce keeps (i, m12[i]) only for m12[i] != -1), so their column direction is
//      skipped: a tiny kernel writes the swapped pair (b, a) for the live pairs and (dummy, dummy) for the
//      others into `pairs_scratch`, where `dummy_image` is an image-table entry with 0 features (its work
//      items are invalid and cost one decode each);
//   3. the same GEMM kernel over `pairs_scratch`, row direction only, with the output pointers advanced to
//      the column-direction halves of mbuf / aux / cand_* -- bit-identical to what direction 1 computes;
//   4. the exact resolve kernel over both directions, as before.
// Results (mbuf of every pair with at least one m12 entry, hence the match lists) are identical to
// launch_k1_filter with n_dirs = 2; non-overlapping pairs cost one GEMM instead of two.
cudaError_t launch_k1_filter_skip(const CUtensorMap& tmap, const MatchParams& p, const uint8_t* desc, int n_pairs,
                                  int max_strips, int num_sms, int32_t* pairs_scratch, int dummy_image,
                                  cudaStream_t stream, cudaEvent_t after_filter);
// Cross-check schedule with a GATHERED column direction (replaces launch_k1_filter_skip): m21 is consulted only at the
// columns some row matched, so the column direction is computed for those columns only:
//   1. GEMM + filter over all pairs, row direction; 2. exact resolve of the row direction (m12);
//   3. per pair the distinct columns j = m12[i] are ranked (ascending), their descriptors copied into a scratch
//      "gathered image" (zero-padded to 256 rows) and one work item per 256 gathered rows appended to a device list
//      (pairs without a match contribute nothing); 4. the same GEMM kernel over that list: rows = gathered
//      descriptors of image b (scratch tensor map), columns = image a -- for a matched column exactly what the
//      full column direction computes; 5. exact resolve of it.  The compaction kernel looks m21 up by rank (colrank).
struct GatherScratch {
  uint8_t* desc;        // [batch x mstride x 128]
  int32_t* colrank;     // [batch][mstride]
  int32_t* cols;        // [batch][mstride] gathered row -> column
  int32_t* cnt;         // [batch]
  int32_t* items;       // [batch x mstride / 256][2]
  int32_t* n_items;     // [1]
};
cudaError_t launch_k1_filter_gather(const CUtensorMap& tmap, const CUtensorMap& tmap_gath, const MatchParams& p,
                                    const uint8_t* desc, int n_pairs, int max_strips, int num_sms, const GatherScratch& g,
                                    cudaStream_t stream, cudaEvent_t after_filter);
cudaError_t launch_k1_gather_phase(int phase, const CUtensorMap& tmap, const CUtensorMap& tmap_gath, const MatchParams& p,
                                   const uint8_t* desc, int n_pairs, int max_strips, int num_sms, const GatherScratch& g,
                                   cudaStream_t stream);
cudaError_t launch_k1_guided(const CUtensorMap& tmap, const MatchParams& p, const GuidedParams& g, int n_pairs,
                             int max_strips, int n_dirs, cudaStream_t stream);
// Cross-check variant: the row direction in full, then the column direction for the matched columns only (the
// gather kernel and scratch of launch_k1_filter_gather); the compaction must look m21 up through g.colrank.
cudaError_t launch_k1_guided_gather(const CUtensorMap& tmap, const CUtensorMap& tmap_gath, const MatchParams& p,
                                    const GuidedParams& g, const uint8_t* desc, int n_pairs, int max_strips,
                                    const GatherScratch& gs, cudaStream_t stream);
// the gather step alone (match_filter_kernel.cu); enable: optional device [n_pairs], pairs with enable < 0 are skipped
cudaError_t launch_gather_matched_columns(const MatchParams& p, const uint8_t* desc, int n_pairs, const GatherScratch& g,
                                          const int32_t* enable, cudaStream_t stream);
cudaError_t launch_crosscheck_compact(const CompactParams& p, int n_pairs, cudaStream_t stream);
// Self-test helper: sets *mismatch != 0 when the match lists of two compaction runs over the same batch
// differ (per pair: count, then every (idx1, idx2) entry; arena offsets may differ between runs).
cudaError_t launch_compare_matches(const uint2* arena_a, const int64_t* off_a, const int32_t* cnt_a, const uint2* arena_b,
                                   const int64_t* off_b, const int32_t* cnt_b, int n_pairs, int32_t* mismatch,
                                   cudaStream_t stream);

}  // namespace b2m
