// match_kernel.cu -- K1: fused all-pairs uint8 descriptor GEMM (tcgen05.mma.kind::i8, TMA-staged
// operands, accumulators in TMEM) + per-row top-2 / lowest-index arg-max / acos-LUT ratio &
// distance tests, and the cross-check + ordered compaction kernel.  The N x M distance matrix
// never leaves the SM.
//
// Semantics follow U:feature/sift.cc (COLMAP 3.9.1) ComputeSiftDistanceMatrix /
// FindBestMatchesOneWayBruteForce / FindBestMatchesBruteForce, reached from
// R:pipeline/match_features.h:45-48.
#include "match_kernel.cuh"
#include "ptx.cuh"

namespace b2m {

namespace {

constexpr int kDim = 128;            // descriptor bytes == K of the GEMM
constexpr int kTileM = 128;          // rows of A per CTA (TMEM lanes)
constexpr int kTileN = 256;          // columns of B per MMA tile
constexpr int kUmmaK = 32;           // bytes of K per tcgen05.mma.kind::i8
constexpr int kStages = 4;           // B-tile ring depth
constexpr int kAccStages = 2;        // TMEM accumulator double buffer (2 x 256 columns = 512)
constexpr int kBytesA = kTileM * kDim;        // 16 KiB
constexpr int kBytesB = kTileN * kDim;        // 32 KiB
constexpr int kEpiWarps = 4;
constexpr int kThreads = (kEpiWarps + 2) * 32;  // 4 epilogue warps + TMA warp + MMA warp
constexpr uint32_t kIdesc = make_idesc_u8u8_s32(kTileM, kTileN);

struct __align__(8) Barriers {
  uint64_t full_a;
  uint64_t full_b[kStages];
  uint64_t empty_b[kStages];
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint32_t tmem_base;
};

constexpr size_t kSmemBytes = 1024 /*align slack*/ + kBytesA + kStages * kBytesB + sizeof(Barriers);

// Merge two tile-local (largest, second-largest) key pairs.  Keys are unique inside a tile.
__device__ __forceinline__ void merge_top2(uint32_t& a1, uint32_t& a2, uint32_t b1, uint32_t b2) {
  const uint32_t lo = min(a1, b1);
  a1 = max(a1, b1);
  a2 = max(max(a2, b2), lo);
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
b2m_k1_match_kernel(const __grid_constant__ CUtensorMap tmap, const MatchParams p) {
  const int pair = blockIdx.z;
  const int dir = blockIdx.y;
  const int strip = blockIdx.x;
  const int ia = p.pairs[2 * pair + dir];
  const int ib = p.pairs[2 * pair + (dir ^ 1)];
  const int nA = p.img_nfeat[ia];
  const int nB = p.img_nfeat[ib];
  if (strip * kTileM >= nA) return;  // uniform exit before any barrier / TMEM allocation
  const int rowA = p.img_row0[ia] + strip * kTileM;
  const int rowB = p.img_row0[ib];
  const int n_tiles = (nB + kTileN - 1) / kTileN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + kBytesA;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kBytesA + kStages * kBytesB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(&bars->full_a, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full_b[s], 1);
      mbar_init(&bars->empty_b[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], kEpiWarps * 32);
    }
    fence_mbar_init();
  }
  if (warp == kEpiWarps + 1) {
    tmem_alloc(&bars->tmem_base, kAccStages * kTileN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kEpiWarps) {
    // ===== TMA producer =====
    if (lane == 0 && n_tiles > 0) {
      mbar_arrive_expect_tx(&bars->full_a, kBytesA);
      tma_load_2d(smA, &tmap, &bars->full_a, 0, rowA);
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&bars->empty_b[stage], phase ^ 1);
        mbar_arrive_expect_tx(&bars->full_b[stage], kBytesB);
        uint8_t* dst = smB + stage * kBytesB;
        tma_load_2d(dst, &tmap, &bars->full_b[stage], 0, rowB + t * kTileN);
        tma_load_2d(dst + kBytesA, &tmap, &bars->full_b[stage], 0, rowB + t * kTileN + 128);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && n_tiles > 0) {
      mbar_wait(&bars->full_a, 0);
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smA));
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        mbar_wait(&bars->full_b[stage], phase);
        tc_fence_after();
        const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smB + stage * kBytesB));
        const uint32_t tmem_d = tmem_base + as * kTileN;
#pragma unroll
        for (int k = 0; k < kDim / kUmmaK; ++k) {
          // +32 bytes of K inside the 128-B swizzle atom == +2 in the (addr >> 4) field
          mma_i8_ss(tmem_d, adesc0 + 2 * k, bdesc0 + 2 * k, kIdesc, k > 0 ? 1u : 0u);
        }
        mma_commit(&bars->empty_b[stage]);  // smem stage reusable once these MMAs retire
        mma_commit(&bars->tmem_full[as]);   // accumulator ready for the epilogue
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
        if (++as == kAccStages) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue: thread <-> row (TMEM lane); running top-2 over all column tiles =====
    const int row_in_strip = warp * 32 + lane;
    int32_t best_d = 0, best_c = -1, second_d = 0;
    uint32_t as = 0, aphase = 0;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + as * kTileN;
      uint32_t k1[4] = {0, 0, 0, 0}, k2[4] = {0, 0, 0, 0};
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr, va);
#pragma unroll
      for (int c = 0; c < kTileN / 32; ++c) {
        tmem_wait_ld();
        uint32_t(&cur)[32] = (c & 1) ? vb : va;
        uint32_t(&nxt)[32] = (c & 1) ? va : vb;
        if (c + 1 < kTileN / 32) tmem_ld_32x32(taddr + (c + 1) * 32, nxt);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          // key = dot * 256 + (255 - local column): max picks the largest dot, lowest column
          const uint32_t key = (cur[j] << 8) | static_cast<uint32_t>(255 - (c * 32 + j));
          const uint32_t lo = min(k1[j & 3], key);
          k1[j & 3] = max(k1[j & 3], key);
          k2[j & 3] = max(k2[j & 3], lo);
        }
      }
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&bars->tmem_empty[as]);
      merge_top2(k1[0], k2[0], k1[1], k2[1]);
      merge_top2(k1[2], k2[2], k1[3], k2[3]);
      merge_top2(k1[0], k2[0], k1[2], k2[2]);
      const int32_t d1 = static_cast<int32_t>(k1[0] >> 8);
      const int32_t d2 = static_cast<int32_t>(k2[0] >> 8);
      if (d1 > best_d) {  // strict: an equal dot in a later tile never displaces an earlier column
        second_d = max(best_d, d2);
        best_d = d1;
        best_c = t * kTileN + (255 - static_cast<int32_t>(k1[0] & 255u));
      } else {
        second_d = max(second_d, d1);
      }
      if (++as == kAccStages) {
        as = 0;
        aphase ^= 1;
      }
    }
    int32_t out = -1;
    if (best_d > 0) {
      const float a = __ldg(p.acos_lut + min(best_d, 262144));
      if (!(a > p.max_distance)) {
        const float b = __ldg(p.acos_lut + min(second_d, 262144));
        if (!(a >= __fmul_rn(p.max_ratio, b))) out = best_c;
      }
    }
    p.mbuf[(static_cast<int64_t>(pair) * 2 + dir) * p.mstride + strip * kTileM + row_in_strip] = out;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAccStages * kTileN);
  }
}

// Cross-check + ordered compaction.  One CTA per pair.  FindBestMatchesBruteForce tail
// (U:feature/sift.cc): keep (i1, m12[i1]) iff m12[i1] != -1 and (no cross-check or
// m21[m12[i1]] == i1); output sorted by i1.
__global__ void __launch_bounds__(256) b2m_crosscheck_compact_kernel(const CompactParams p) {
  const int pair = blockIdx.x;
  const int i1 = p.pairs[2 * pair];
  const int n1 = p.img_nfeat[i1];
  const int32_t* m12 = p.mbuf + (static_cast<int64_t>(pair) * 2) * p.mstride;
  const int32_t* m21 = m12 + p.mstride;
  // gathered column direction: m21 holds one entry per MATCHED column, at the column's rank
  const int32_t* rank = p.colrank ? p.colrank + static_cast<int64_t>(pair) * p.mstride : nullptr;
  auto mutual = [&](int i, int j) { return m21[rank ? rank[j] : j] == i; };
  __shared__ int s_warp[8];
  __shared__ int s_base;
  __shared__ unsigned long long s_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.enable && p.enable[pair] < 0) {  // guided pass: this pair keeps its verified inliers
    if (threadIdx.x == 0) {
      p.pair_off[pair] = 0;
      p.pair_cnt[pair] = -1;
    }
    return;
  }

  if (p.cand_cnt && p.cand_cnt[2 * pair] == 0) {  // most pairs of an exhaustive batch: nothing matched
    if (threadIdx.x == 0) {
      p.pair_off[pair] = static_cast<int64_t>(atomicAdd(p.cursor, 0ull));
      p.pair_cnt[pair] = 0;
    }
    return;
  }

  // pass 1: count
  int cnt = 0;
  for (int i = threadIdx.x; i < n1; i += 256) {
    const int j = m12[i];
    cnt += (j >= 0 && (!p.cross_check || mutual(i, j))) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_warp[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < 8; ++w) total += s_warp[w];
    s_off = atomicAdd(p.cursor, static_cast<unsigned long long>(total));
    p.pair_off[pair] = static_cast<int64_t>(s_off);
    p.pair_cnt[pair] = total;
    s_base = 0;
    s_warp[0] = total;
  }
  __syncthreads();
  if (s_warp[0] == 0) return;   // uniform: nothing to write
  __syncthreads();              // s_warp is reused by the ordered write
  uint2* out = p.arena + s_off;

  // pass 2: ordered write
  for (int i0 = 0; i0 < n1; i0 += 256) {
    const int i = i0 + threadIdx.x;
    int j = -1;
    bool keep = false;
    if (i < n1) {
      j = m12[i];
      keep = (j >= 0 && (!p.cross_check || mutual(i, j)));
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int wpre = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int base = s_base;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    if (keep) {
      out[base + wpre] = make_uint2(static_cast<unsigned>(i), static_cast<unsigned>(j));
      if (p.pts) {  // matched pixel coordinates for the verifier (keypoints are float32, exact in double)
        const float2 a = p.kpts[p.img_row0[i1] + i];
        const float2 b = p.kpts[p.img_row0[p.pairs[2 * pair + 1]] + j];
        p.pts[s_off + base + wpre] = make_double4(a.x, a.y, b.x, b.y);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int total = 0;
      for (int w = 0; w < 8; ++w) total += s_warp[w];
      s_base += total;
    }
    __syncthreads();
  }
}

cudaError_t launch_k1_match(const CUtensorMap& tmap, const MatchParams& p, int n_pairs, int max_strips, int n_dirs,
                            cudaStream_t stream) {
  // function attributes are per device: several contexts on different GPUs may live in one process
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(b2m_k1_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  dim3 grid(max_strips, n_dirs, n_pairs);
  b2m_k1_match_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmap, p);
  return cudaGetLastError();
}

cudaError_t launch_crosscheck_compact(const CompactParams& p, int n_pairs, cudaStream_t stream) {
  b2m_crosscheck_compact_kernel<<<n_pairs, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b2m
