/* oracle_match.c -- CPU restatement of the reference's brute-force SIFT matcher.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pycolmap_b200/ may link, import or call this;
 * it is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * PARITY UNPINNED: the reference (/root/reference, pycolmap bindings) forwards into COLMAP
 * 3.9.1 (R:CMakeLists.txt:17, R:pyproject.toml:36), which is not vendored and has no golden
 * vectors for this path.  This file restates the published algorithm
 * of U:feature/sift.cc (COLMAP 3.9.1):
 *     ComputeSiftDistanceMatrix          -> orc_dist_matrix        (row M1)
 *     FindBestMatchesOneWayBruteForce    -> orc_best_one_way       (row M2)
 *     FindBestMatchesBruteForce          -> orc_match_bruteforce   (row M3)
 *     MatchGuidedSiftFeaturesCPU         -> orc_match_guided       (row G1)
 * anchored on the reference call sites R:pipeline/match_features.h:45-48 (matcher factory ->
 * FeatureMatcherWorker -> Match) and the PyFeatureMatches layout R:estimators/two_view_geometry.h:19-38.
 *
 * Two implementations, asserted identical by tests/test_oracle_match.py:
 *   (1) the literal one: materialise the n1 x n2 int32 matrix, scan rows, scan columns;
 *   (2) orc_fast_*: streaming, SIMD (AVX-512 VNNI / AVX2 / generic, picked at run time),
 *       threaded over image pairs like upstream's FeatureMatcherWorker pool -- the timed
 *       CPU baseline.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#define DIM 128

/* ---- (1) literal restatement ------------------------------------------------------- */

/* M1: dists(i1,i2) = sum_k d1[i1,k] * d2[i2,k], int32 (uint8 promoted to int). */
void orc_dist_matrix(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int32_t* dists) {
  for (int i1 = 0; i1 < n1; ++i1) {
    const uint8_t* a = d1 + (size_t)i1 * DIM;
    for (int i2 = 0; i2 < n2; ++i2) {
      const uint8_t* b = d2 + (size_t)i2 * DIM;
      int32_t s = 0;
      for (int k = 0; k < DIM; ++k) s += (int32_t)a[k] * (int32_t)b[k];
      dists[(size_t)i1 * n2 + i2] = s;
    }
  }
}

/* The float32 accept test of M2, shared by every implementation in this file.
 * best / second are the integer dot products; returns 1 when the match is kept. */
static inline int accept_match(int32_t best, int32_t second, float max_ratio, float max_distance) {
  const float kDistNorm = 1.0f / (512.0f * 512.0f);
  const float best_dist_normed = acosf(fminf(kDistNorm * (float)best, 1.0f));
  if (best_dist_normed > max_distance) return 0;
  const float second_best_dist_normed = acosf(fminf(kDistNorm * (float)second, 1.0f));
  volatile float rhs = max_ratio * second_best_dist_normed; /* volatile: no FMA contraction */
  if (best_dist_normed >= rhs) return 0;
  return 1;
}

/* M2 on a strided view: element (r, c) is dists[r*stride_r + c*stride_c]. */
void orc_best_one_way(const int32_t* dists, int rows, int cols, int64_t stride_r, int64_t stride_c,
                      float max_ratio, float max_distance, int32_t* matches) {
  for (int r = 0; r < rows; ++r) {
    int32_t best_c = -1;
    int32_t best = 0, second = 0;
    for (int c = 0; c < cols; ++c) {
      const int32_t d = dists[r * stride_r + c * stride_c];
      if (d > best) {
        best_c = c;
        second = best;
        best = d;
      } else if (d > second) {
        second = d;
      }
    }
    matches[r] = -1;
    if (best_c == -1) continue;
    if (!accept_match(best, second, max_ratio, max_distance)) continue;
    matches[r] = best_c;
  }
}

/* M3 given a distance matrix.  out_matches: [<=min(n1,n2)] x (idx1, idx2).  Returns count. */
int orc_matches_from_matrix(const int32_t* dists, int n1, int n2, float max_ratio, float max_distance,
                            int cross_check, uint32_t* out_matches) {
  int32_t* m12 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n1 > 0 ? n1 : 1));
  int32_t* m21 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
  orc_best_one_way(dists, n1, n2, n2, 1, max_ratio, max_distance, m12);
  if (cross_check) orc_best_one_way(dists, n2, n1, 1, n2, max_ratio, max_distance, m21);
  int cnt = 0;
  for (int i1 = 0; i1 < n1; ++i1) {
    const int32_t j = m12[i1];
    if (j == -1) continue;
    if (cross_check && m21[j] != i1) continue;
    out_matches[2 * cnt + 0] = (uint32_t)i1;
    out_matches[2 * cnt + 1] = (uint32_t)j;
    ++cnt;
  }
  free(m12);
  free(m21);
  return cnt;
}

/* M3: full literal pipeline. */
int orc_match_bruteforce(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float max_ratio,
                         float max_distance, int cross_check, uint32_t* out_matches) {
  if (n1 <= 0 || n2 <= 0) return 0;
  int32_t* dists = (int32_t*)malloc(sizeof(int32_t) * (size_t)n1 * (size_t)n2);
  if (!dists) return -1;
  orc_dist_matrix(d1, n1, d2, n2, dists);
  int cnt = orc_matches_from_matrix(dists, n1, n2, max_ratio, max_distance, cross_check, out_matches);
  free(dists);
  return cnt;
}

/* G1: guided matching.  kind: 0 = F/E (squared Sampson, float32), 1 = H (forward transfer).
 * model: 9 floats row-major.  kp: [n x 2] float32 (x, y).  Entries failing the geometric test
 * get dist = 0 before M2/M3. */
int orc_match_guided(const uint8_t* d1, const float* kp1, int n1, const uint8_t* d2, const float* kp2,
                     int n2, int kind, const float* M, float max_error, float max_ratio,
                     float max_distance, int cross_check, uint32_t* out_matches) {
  if (n1 <= 0 || n2 <= 0) return 0;
  int32_t* dists = (int32_t*)malloc(sizeof(int32_t) * (size_t)n1 * (size_t)n2);
  if (!dists) return -1;
  orc_dist_matrix(d1, n1, d2, n2, dists);
  const float thr = max_error * max_error;
  for (int i1 = 0; i1 < n1; ++i1) {
    const float x1 = kp1[2 * i1], y1 = kp1[2 * i1 + 1];
    for (int i2 = 0; i2 < n2; ++i2) {
      const float x2 = kp2[2 * i2], y2 = kp2[2 * i2 + 1];
      float r;
      if (kind == 0) {
        const float Fx0 = M[0] * x1 + M[1] * y1 + M[2];
        const float Fx1 = M[3] * x1 + M[4] * y1 + M[5];
        const float Fx2 = M[6] * x1 + M[7] * y1 + M[8];
        const float Ft0 = M[0] * x2 + M[3] * y2 + M[6];
        const float Ft1 = M[1] * x2 + M[4] * y2 + M[7];
        const float num = x2 * Fx0 + y2 * Fx1 + Fx2;
        r = num * num / (Fx0 * Fx0 + Fx1 * Fx1 + Ft0 * Ft0 + Ft1 * Ft1);
      } else {
        const float w = M[6] * x1 + M[7] * y1 + M[8];
        const float u = (M[0] * x1 + M[1] * y1 + M[2]) / w;
        const float v = (M[3] * x1 + M[4] * y1 + M[5]) / w;
        r = (u - x2) * (u - x2) + (v - y2) * (v - y2);
      }
      if (!(r <= thr)) dists[(size_t)i1 * n2 + i2] = 0;
    }
  }
  int cnt = orc_matches_from_matrix(dists, n1, n2, max_ratio, max_distance, cross_check, out_matches);
  free(dists);
  return cnt;
}

/* ---- (2) streaming SIMD implementation (timed CPU baseline) ------------------------ */

typedef void (*dot_row_fn)(const uint8_t* a, const uint8_t* b, int n2, int32_t* out);

static void dot_row_generic(const uint8_t* a, const uint8_t* b, int n2, int32_t* out) {
  for (int j = 0; j < n2; ++j) {
    const uint8_t* bj = b + (size_t)j * DIM;
    int32_t s = 0;
    for (int k = 0; k < DIM; ++k) s += (int32_t)a[k] * (int32_t)bj[k];
    out[j] = s;
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void dot_row_avx2(const uint8_t* a, const uint8_t* b, int n2,
                                                         int32_t* out) {
  __m256i a16[8];
  for (int t = 0; t < 8; ++t)
    a16[t] = _mm256_cvtepu8_epi16(_mm_loadu_si128((const __m128i*)(a + 16 * t)));
  for (int j = 0; j < n2; ++j) {
    const uint8_t* bj = b + (size_t)j * DIM;
    __m256i acc = _mm256_setzero_si256();
    for (int t = 0; t < 8; ++t) {
      __m256i b16 = _mm256_cvtepu8_epi16(_mm_loadu_si128((const __m128i*)(bj + 16 * t)));
      acc = _mm256_add_epi32(acc, _mm256_madd_epi16(a16[t], b16));
    }
    __m128i s = _mm_add_epi32(_mm256_castsi256_si128(acc), _mm256_extracti128_si256(acc, 1));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0x4E));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0xB1));
    out[j] = _mm_cvtsi128_si32(s);
  }
}

/* u8 x u8 through vpdpbusd (u8 x s8): a.b = a.(b-128) + 128*sum(a); b-128 == b ^ 0x80 as s8. */
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vnni"))) static void dot_row_vnni(
    const uint8_t* a, const uint8_t* b, int n2, int32_t* out) {
  const __m512i a0 = _mm512_loadu_si512((const void*)a);
  const __m512i a1 = _mm512_loadu_si512((const void*)(a + 64));
  int32_t asum = 0;
  for (int k = 0; k < DIM; ++k) asum += a[k];
  const int32_t bias = 128 * asum;
  const __m512i flip = _mm512_set1_epi8((char)0x80);
  int j = 0;
  for (; j + 4 <= n2; j += 4) {
    __m512i acc[4];
    for (int u = 0; u < 4; ++u) {
      const uint8_t* bj = b + (size_t)(j + u) * DIM;
      __m512i b0 = _mm512_xor_si512(_mm512_loadu_si512((const void*)bj), flip);
      __m512i b1 = _mm512_xor_si512(_mm512_loadu_si512((const void*)(bj + 64)), flip);
      __m512i c = _mm512_dpbusd_epi32(_mm512_setzero_si512(), a0, b0);
      acc[u] = _mm512_dpbusd_epi32(c, a1, b1);
    }
    for (int u = 0; u < 4; ++u) out[j + u] = _mm512_reduce_add_epi32(acc[u]) + bias;
  }
  for (; j < n2; ++j) {
    const uint8_t* bj = b + (size_t)j * DIM;
    __m512i b0 = _mm512_xor_si512(_mm512_loadu_si512((const void*)bj), flip);
    __m512i b1 = _mm512_xor_si512(_mm512_loadu_si512((const void*)(bj + 64)), flip);
    __m512i c = _mm512_dpbusd_epi32(_mm512_setzero_si512(), a0, b0);
    c = _mm512_dpbusd_epi32(c, a1, b1);
    out[j] = _mm512_reduce_add_epi32(c) + bias;
  }
}
#endif

static dot_row_fn pick_dot_row(const char** name) {
#if defined(__x86_64__)
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512vnni") && __builtin_cpu_supports("avx512bw") &&
      __builtin_cpu_supports("avx512vl")) {
    if (name) *name = "avx512vnni";
    return dot_row_vnni;
  }
  if (__builtin_cpu_supports("avx2")) {
    if (name) *name = "avx2";
    return dot_row_avx2;
  }
#endif
  if (name) *name = "generic";
  return dot_row_generic;
}

const char* orc_fast_isa(void) {
  const char* n = "generic";
  pick_dot_row(&n);
  return n;
}

/* Streaming M1+M2+M3 for one pair; same scan order as the literal version so tie-breaking
 * is identical: rows ascending outer, columns ascending inner. */
int orc_fast_match_pair(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float max_ratio,
                        float max_distance, int cross_check, uint32_t* out_matches) {
  if (n1 <= 0 || n2 <= 0) return 0;
  dot_row_fn dot_row = pick_dot_row(NULL);
  int32_t* row = (int32_t*)malloc(sizeof(int32_t) * (size_t)n2);
  int32_t* m12 = (int32_t*)malloc(sizeof(int32_t) * (size_t)n1);
  int32_t* cbest = (int32_t*)calloc((size_t)n2, sizeof(int32_t));
  int32_t* csecond = (int32_t*)calloc((size_t)n2, sizeof(int32_t));
  int32_t* cidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n2);
  for (int j = 0; j < n2; ++j) cidx[j] = -1;
  for (int i = 0; i < n1; ++i) {
    dot_row(d1 + (size_t)i * DIM, d2, n2, row);
    int32_t best = 0, second = 0, best_j = -1;
    for (int j = 0; j < n2; ++j) {
      const int32_t d = row[j];
      if (d > best) {
        best_j = j;
        second = best;
        best = d;
      } else if (d > second) {
        second = d;
      }
    }
    m12[i] = (best_j != -1 && accept_match(best, second, max_ratio, max_distance)) ? best_j : -1;
    if (cross_check) {
      for (int j = 0; j < n2; ++j) {
        const int32_t d = row[j];
        if (d > cbest[j]) {
          cidx[j] = i;
          csecond[j] = cbest[j];
          cbest[j] = d;
        } else if (d > csecond[j]) {
          csecond[j] = d;
        }
      }
    }
  }
  int cnt = 0;
  for (int i = 0; i < n1; ++i) {
    const int32_t j = m12[i];
    if (j == -1) continue;
    if (cross_check) {
      if (cidx[j] != i) continue;
      if (!accept_match(cbest[j], csecond[j], max_ratio, max_distance)) continue;
    }
    out_matches[2 * cnt] = (uint32_t)i;
    out_matches[2 * cnt + 1] = (uint32_t)j;
    ++cnt;
  }
  free(row);
  free(m12);
  free(cbest);
  free(csecond);
  free(cidx);
  return cnt;
}

/* Threaded batch over image pairs (one pair = one unit of work, like upstream's
 * FeatureMatcherWorker pool, U:controllers/feature_matching_utils.cc).
 * desc: packed [sum n_feat x 128]; offsets[i] = first row of image i.
 * out_matches: [n_pairs x stride x 2]; out_counts[n_pairs]. */
typedef struct {
  const uint8_t* desc;
  const int64_t* offsets;
  const int32_t* n_feat;
  const int32_t* pairs;
  int64_t n_pairs;
  float max_ratio, max_distance;
  int cross_check;
  uint32_t* out_matches;
  int64_t stride;
  int32_t* out_counts;
  volatile int64_t* next;
} batch_args;

static void* batch_worker(void* p) {
  batch_args* a = (batch_args*)p;
  for (;;) {
    int64_t k = __sync_fetch_and_add(a->next, 1);
    if (k >= a->n_pairs) break;
    const int i = a->pairs[2 * k], j = a->pairs[2 * k + 1];
    a->out_counts[k] = orc_fast_match_pair(a->desc + a->offsets[i] * DIM, a->n_feat[i],
                                           a->desc + a->offsets[j] * DIM, a->n_feat[j], a->max_ratio,
                                           a->max_distance, a->cross_check,
                                           a->out_matches + (size_t)k * a->stride * 2);
  }
  return NULL;
}

int orc_fast_match_pairs(const uint8_t* desc, const int64_t* offsets, const int32_t* n_feat,
                         const int32_t* pairs, int64_t n_pairs, float max_ratio, float max_distance,
                         int cross_check, int n_threads, uint32_t* out_matches, int64_t stride,
                         int32_t* out_counts) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  volatile int64_t next = 0;
  batch_args a = {desc, offsets, n_feat, pairs, n_pairs, max_ratio, max_distance, cross_check,
                  out_matches, stride, out_counts, &next};
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, batch_worker, &a);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* acosf table used to cross-check the product's device LUT in tests: lut[d] for d in [0, 2^18]. */
void orc_acos_lut(float* lut) {
  const float kDistNorm = 1.0f / (512.0f * 512.0f);
  for (int d = 0; d <= 262144; ++d) lut[d] = acosf(fminf(kDistNorm * (float)d, 1.0f));
}
