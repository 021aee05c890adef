// match_filter_kernel.cu -- K1 (v14): persistent, clustered, pair-MMA tcgen05 int8 GEMM with a *filter*
// epilogue, plus the exact resolve kernel.
//
// Grid = one 2-CTA cluster per SM pair (74 clusters), persistent: every cluster walks a static list of
// work items (image pair, direction, 256-row block of the "row" image A).
//   * The two CTAs form a cta_group::2 pair: ONE elected thread of the leader CTA issues
//     tcgen05.mma.cta_group::2.kind::i8 256x256x32 (four k-steps per 256-column tile of image B).  Each SM
//     contributes its own 128-row A strip and stages only ITS 128-column half of every B tile, so each
//     operand byte is written to and read from shared memory once per pair of SMs.
//   * B streams through an 8-stage ring of 16 KiB half tiles (TMA, SWIZZLE_128B, both halves accounted on
//     the leader's mbarrier); the A strips are double-buffered across work items, so the pipelines
//     (TMA ring, TMEM accumulator stages, mbarrier phases) run straight across item boundaries and
//     TMEM / barriers are set up once per launch.
//   * Accumulators: TMEM, 2 stages x 256 columns = all 512 columns of each SM; every SM drains its own
//     128 lanes.  Stage hand-back to the leader = one cluster-scope mbarrier arrival per epilogue warp
//     (aggregating them on a shared-memory counter first was measured slower).
//   * The TMA and MMA warps stay warp-converged and predicate only the tcgen05 / TMA instructions with
//     elect.sync, so descriptors and addresses live in uniform registers.
//
// Warp roles (19 warps, <= 96 registers per thread):
//   warps 0..15  epilogue: warp w drains TMEM lane quarter w%4, column group w/4 (64 columns) of every tile
//                with ONE round trip (four tcgen05.ld.x16 in flight, one wait::ld), hands the stage back, then
//                folds the 64 accumulators of its row into 16 slot maxima.  What bounds K1 is the hand-shake
//                chain MMA -> commit -> drain -> arrive -> MMA over only two accumulator stages, so the drain
//                is kept to a single TMEM round trip and nothing else sits between wait::ld and the arrival.
//   warp 16      TMA producer;  warp 17  MMA issuer (leader CTA only);
//   warp 18      selector: turns the slot maxima of a finished item into reject / candidate decisions one item
//                behind the epilogue warps, so the dependent global latencies (acos table, candidate counter)
//                are off the chain.
//
// Filter: instead of an exact running top-2 (4 ALU ops per accumulator) each row keeps 64 "slot maxima"
//     slot(g, r) = max over columns j = 256 t + 64 g + 16 c + r   (g = 0..3, r = 0..15; c = 0..3, all tiles t)
// updated with two 3-input max (VIMNMX3) per four accumulators = 0.5 ALU op per accumulator.  At the end
// of the row
//     best = max over slots (exact);   S1 = second largest slot maximum (multiset), which is a LOWER
//     bound of the true second-best (= max(S1, second largest element inside the winning slot)).
// acos is monotone, so a row failing `acos(best) <= max_distance`, or failing the ratio test already
// against S1, is rejected exactly.  The survivors ("candidates": essentially the true matches) are
// resolved exactly by b2m_k1_resolve_kernel: it recomputes the n2/64 dot products of the winning slot
// with dp4a (all columns if several slots share the maximum), finds the lowest-index arg-max and the
// hidden second-best, and applies the float32 test of FindBestMatchesOneWayBruteForce.  The match
// indices are bit-identical to the exact kernel (match_kernel.cu) and to the CPU oracle.
//
// -DB2M_K1_PROF adds clock64 role counters (printed every 8th launch); see DESIGN.md for the readings.
//
// Semantics: U:feature/sift.cc (COLMAP 3.9.1), SURVEY.md section 8 rows M1-M3.
#include <cstdio>
#include "match_kernel.cuh"
#include "ptx.cuh"

namespace b2m {

namespace {

constexpr int kDim = 128;
constexpr int kTileM = 128;                        // rows per CTA and MMA (TMEM lanes)
constexpr int kCluster = 2;                        // CTAs per cluster = one cta_group::2 pair
constexpr int kRowsPerItem = kTileM * kCluster;    // 256 rows of A per work item == kRowPad
constexpr int kTileN = 256;                        // columns per B tile: N = 256 hides the smem A read
                                                   // (N <= 128 costs ~92 cycles per MMA regardless of N, profiles/r01_microbench2)
constexpr int kUmmaK = 32;
constexpr int kStages = 8;                         // B-tile ring depth (8 x 16 KiB: each CTA stages only ITS half)
constexpr int kAccStages = 2;
constexpr int kABufs = 2;                          // A strips double-buffered across work items
constexpr int kBytesA = kTileM * kDim;             // 16 KiB per strip
constexpr int kBytesB = (kTileN / kCluster) * kDim;  // 16 KiB: this CTA's half (128 columns) of a B tile
constexpr int kEpiWarps = 16;                      // warp w: TMEM lane quarter w%4, column group w/4 of every tile
constexpr int kColGroups = kEpiWarps / 4;          // 4
constexpr int kGroupCols = kTileN / kColGroups;    // 64 columns per warp and tile = four x16 loads
constexpr int kOwnSlots = 16;                      // slot maxima per row kept in one thread's registers
constexpr int kSlots = kOwnSlots * kColGroups;     // 64 slot maxima per row
constexpr int kSlotCols = kTileN / kSlots;         // 4 columns of every tile share a slot
constexpr int kThreads = (kEpiWarps + 3) * 32;     // + TMA warp + MMA warp + selector warp
constexpr int kAccCols = kTileN;                   // TMEM columns per accumulator stage
constexpr uint32_t kIdesc = make_idesc_u8u8_s32(kTileM * kCluster, kTileN);  // one 256 x 256 x 32 MMA per CTA pair
constexpr uint16_t kClusterMask = static_cast<uint16_t>((1u << kCluster) - 1u);
static_assert(kRowsPerItem == kRowPad && kTileN == kRowPad, "images are padded to whole items / column tiles");
static_assert(kGroupCols == 64 && kSlotCols == 4, "epilogue folds four x16 chunks into 16 slots");

struct __align__(8) Barriers {
  uint64_t full_a[kABufs];
  uint64_t empty_a[kABufs];
  uint64_t full_b[kStages];
  uint64_t empty_b[kStages];
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint64_t sel_full, sel_empty;  // slot maxima of an item handed to / consumed by the selector warp
  uint32_t tmem_base;
};

constexpr int kMergeBytes = kTileM * kSlots * 4;   // slot maxima of one item, [slot][128 rows]
constexpr size_t kSmemBytes = 1024 + kABufs * kBytesA + kStages * kBytesB + kMergeBytes + sizeof(Barriers);

// A work item and the data every warp role derives from its index (pure function of `w`).
struct Item {
  int pair, dir, nA, nB, rowA, rowB, n_tiles, row0;  // row0: first row (inside image A) of this CTA
  bool valid;
};
__device__ __forceinline__ Item decode_item(const MatchParams& p, int w, uint32_t cta_rank) {
  Item it;
  if (p.item_list) {
    // gathered column direction: rows = the gathered block of the pair (scratch tensor map), columns = image a
    it.pair = p.item_list[2 * w];
    const int cb = p.item_list[2 * w + 1];
    it.dir = 0;
    const int ib = p.pairs[2 * it.pair];
    it.nA = p.gath_cnt[it.pair];
    it.nB = p.img_nfeat[ib];
    it.valid = cb * kRowsPerItem < it.nA;
    it.row0 = cb * kRowsPerItem + static_cast<int>(cta_rank) * kTileM;
    it.rowA = it.pair * p.mstride + it.row0;
    it.rowB = p.img_row0[ib];
    it.n_tiles = (it.nB + kTileN - 1) / kTileN;
    return it;
  }
  const int cb = w % p.blocks_per_image;
  const int pd = w / p.blocks_per_image;
  it.dir = pd % p.n_dirs;
  it.pair = pd / p.n_dirs;
  const int ia = p.pairs[2 * it.pair + it.dir];
  const int ib = p.pairs[2 * it.pair + (it.dir ^ 1)];
  it.nA = p.img_nfeat[ia];
  it.nB = p.img_nfeat[ib];
  // same answer in both CTAs of the cluster; an empty image B still yields an item (n_tiles == 0)
  // so that its rows are written as "no match"
  it.valid = cb * kRowsPerItem < it.nA;
  it.row0 = cb * kRowsPerItem + static_cast<int>(cta_rank) * kTileM;
  it.rowA = p.img_row0[ia] + it.row0;
  it.rowB = p.img_row0[ib];
  it.n_tiles = (it.nB + kTileN - 1) / kTileN;
  return it;
}

}  // namespace

#ifdef B2M_K1_PROF
// role counters (SM cycles, summed over CTAs / warps): [0] MMA wait tmem_empty, [1] MMA wait full_b, [2] MMA issue,
// [3] MMA tiles, [4] epi wait tmem_full, [5] epi drain (first ld .. wait::ld), [6] epi hand-back + fold,
// [7] epi warp-tiles, [8] epi item tail (slot maxima -> shared memory), [9] epi items
__device__ unsigned long long g_k1_prof[16];
#define PROF_T(x) const long long x = clock64()
#define PROF_ADD(acc, a, b) acc += (b) - (a)
#else
#define PROF_T(x)
#define PROF_ADD(acc, a, b)
#endif

// elect.sync: true in exactly one (converged) lane.  The MMA / TMA roles keep their whole warp converged
// and predicate only the tcgen05 / TMA instructions, so descriptors and addresses stay warp-uniform
// (uniform registers) instead of being rebuilt in vector registers and moved over for every instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
b2m_k1_filter_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_a,
                     const MatchParams p) {
  // tmap: the resident descriptor set (column tiles; row strips too unless the rows are gathered), tmap_a: the row
  // strips (== tmap except for the gathered column direction, where it covers the gather scratch)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                                     // [kABufs][16 KiB]
  uint8_t* smB = smem + kABufs * kBytesA;                  // [kStages][16 KiB]
  uint32_t* merge = reinterpret_cast<uint32_t*>(smB + kStages * kBytesB);
  Barriers* bars = reinterpret_cast<Barriers*>(smB + kStages * kBytesB + kMergeBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / kCluster;
  const int n_clusters = gridDim.x / kCluster;
  // gathered column direction: the work list was built on the device, so was its length
  const int n_items = p.item_list ? __ldg(p.n_items_ptr) : p.n_items;

  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&tmap);
    tma_prefetch_desc(&tmap_a);
    for (int s = 0; s < kABufs; ++s) {
      mbar_init(&bars->full_a[s], 1);
      mbar_init(&bars->empty_a[s], 1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full_b[s], 1);
      mbar_init(&bars->empty_b[s], 1);  // the leader's pair-commit arrives here in both CTAs
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], kCluster * kEpiWarps);  // leader only: one arrival per epilogue warp of the pair
    }
    mbar_init(&bars->sel_full, kEpiWarps);
    mbar_init(&bars->sel_empty, 1);
    fence_mbar_init();
  }
  if (warp == kEpiWarps + 1) {
    tmem_alloc_pair(&bars->tmem_base, kAccStages * kAccCols);  // executed by the same warp of both CTAs
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before any remote arrival / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kEpiWarps) {
    // ===== TMA producer (whole warp converged, one elected lane issues) =====
    uint32_t stage = 0, phase = 0, n_done = 0;
    for (int w = cluster_id; w < n_items; w += n_clusters) {
      const Item it = decode_item(p, w, cta_rank);
      if (!it.valid) continue;
      const uint32_t ab = n_done & 1, aph = (n_done >> 1) & 1;
      ++n_done;
      mbar_wait(&bars->empty_a[ab], aph ^ 1);  // the MMAs of the item that used this A buffer retired
      if (elect_one()) {
        // both CTAs' bytes are accounted on the LEADER's barriers (the leader issues the pair MMA)
        if (cta_rank == 0) mbar_arrive_expect_tx(&bars->full_a[ab], kCluster * kBytesA);
        tma_load_2d_pair(smA + ab * kBytesA, &tmap_a, &bars->full_a[ab], 0, it.rowA);
      }
      __syncwarp();
      for (int t = 0; t < it.n_tiles; ++t) {
        mbar_wait(&bars->empty_b[stage], phase ^ 1);
        if (elect_one()) {
          if (cta_rank == 0) mbar_arrive_expect_tx(&bars->full_b[stage], kCluster * kBytesB);
          // this CTA stages only ITS 128-column half of the tile; the pair MMA reads the other half from
          // the peer's shared memory, so each B byte is written to and read from shared memory once per pair
          tma_load_2d_pair(smB + stage * kBytesB, &tmap, &bars->full_b[stage], 0,
                           it.rowB + t * kTileN + static_cast<int>(cta_rank) * (kTileN / kCluster));
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===== MMA issuer: the leader CTA's warp drives the 256-row MMAs of the pair (one elected lane issues) =====
    if (cta_rank == 0) {
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0, n_done = 0;
      const uint32_t smA_u32 = smem_u32(smA), smB_u32 = smem_u32(smB);
#ifdef B2M_K1_PROF
      long long pm_empty = 0, pm_fullb = 0, pm_issue = 0, pm_tiles = 0;
#endif
      for (int w = cluster_id; w < n_items; w += n_clusters) {
        const Item it = decode_item(p, w, cta_rank);
        if (!it.valid) continue;
        const uint32_t ab = n_done & 1, aph = (n_done >> 1) & 1;
        ++n_done;
        mbar_wait(&bars->full_a[ab], aph);
        const uint64_t adesc = make_smem_desc_sw128(smA_u32 + ab * kBytesA);
        if (it.n_tiles == 0) {  // nothing will read this A buffer: release it in both CTAs
          if (elect_one()) {
            mbar_arrive_cluster(&bars->empty_a[ab], 0);
            mbar_arrive_cluster(&bars->empty_a[ab], 1);
          }
          __syncwarp();
        }
        for (int t = 0; t < it.n_tiles; ++t) {
          // B is normally resident long before the accumulator stage comes back: test it first so that
          // nothing but the descriptor set-up sits between the stage hand-back and the first MMA
          PROF_T(t1);
          mbar_wait(&bars->full_b[stage], phase);
          PROF_T(t2);
          mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
          PROF_T(t0);
          tc_fence_after();
          const uint64_t bdesc = make_smem_desc_sw128(smB_u32 + stage * kBytesB);
          const uint32_t tmem_d = tmem_base + as * kAccCols;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kDim / kUmmaK; ++k)
              mma_i8_ss_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, kIdesc, k > 0 ? 1u : 0u);
            mma_commit_pair(&bars->tmem_full[as], kClusterMask);   // both CTAs' epilogues may drain (critical chain first)
            mma_commit_pair(&bars->empty_b[stage], kClusterMask);  // both CTAs' producers may refill
            if (t == it.n_tiles - 1) mma_commit_pair(&bars->empty_a[ab], kClusterMask);  // A buffers reusable
          }
          __syncwarp();
#ifdef B2M_K1_PROF
          {
            PROF_T(t3);
            PROF_ADD(pm_fullb, t1, t2);
            PROF_ADD(pm_empty, t2, t0);
            PROF_ADD(pm_issue, t0, t3);
            ++pm_tiles;
          }
#endif
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
#ifdef B2M_K1_PROF
      if (lane == 0) {
        atomicAdd(&g_k1_prof[0], (unsigned long long)pm_empty);
        atomicAdd(&g_k1_prof[1], (unsigned long long)pm_fullb);
        atomicAdd(&g_k1_prof[2], (unsigned long long)pm_issue);
        atomicAdd(&g_k1_prof[3], (unsigned long long)pm_tiles);
      }
#endif
    }
  } else if (warp == kEpiWarps + 2) {
    // ===== selector: slot maxima of one item (shared memory, [slot][row]) -> reject / candidate decisions =====
    // Runs one item behind the epilogue warps, so the dependent global-memory latencies of the decision
    // (acos table, candidate counter) are off the accumulator hand-shake chain.  Lane l owns rows l, l+32, ...
    uint32_t sphase = 0;
    for (int w = cluster_id; w < n_items; w += n_clusters) {
      const Item it = decode_item(p, w, cta_rank);
      if (!it.valid) continue;
      mbar_wait(&bars->sel_full, sphase);
      uint32_t best[4], s1[4];
      int sstar[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        best[q] = 0u;
        s1[q] = 0u;
        sstar[q] = 0;
      }
      // (largest, second largest) over the slot maxima, multiset semantics; lowest slot id on ties
#pragma unroll 8
      for (int r = 0; r < kSlots; ++r) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t v = merge[r * kTileM + q * 32 + lane];
          if (v > best[q]) {
            s1[q] = best[q];
            best[q] = v;
            sstar[q] = r;
          } else {
            s1[q] = max(s1[q], v);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->sel_empty);  // the epilogue warps may overwrite the buffer
      sphase ^= 1;
      float fa[4], fb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        fa[q] = __ldg(p.acos_lut + min(best[q], 262144u));
        fb[q] = __ldg(p.acos_lut + min(s1[q], 262144u));
      }
      const int64_t base = (static_cast<int64_t>(it.pair) * 2 + it.dir) * p.mstride;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int32_t out = -1;
        if (best[q] > 0u && !(fa[q] > p.max_distance) && !(fa[q] >= __fmul_rn(p.max_ratio, fb[q])))
          out = -2 - sstar[q];  // candidate: resolve exactly
        const int row = it.row0 + q * 32 + lane;
        p.mbuf[base + row] = out;
        if (out != -1 && row < it.nA) {
          p.aux[base + row] = make_uint2(best[q], s1[q]);
          const int k = atomicAdd(p.cand_cnt + it.pair * 2 + it.dir, 1);
          p.cand_rows[base + k] = row | (sstar[q] << 24);  // winning slot travels with the row
        }
      }
    }
  } else {
    // ===== filter epilogue =====
    const int quarter = warp & 3;
    const int group = warp >> 2;
    const int row_in_cta = quarter * 32 + lane;
    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + group * kGroupCols;
    uint32_t as = 0, aphase = 0, sel_phase = 0;
#ifdef B2M_K1_PROF
    long long pe_full = 0, pe_drain = 0, pe_rest = 0, pe_tiles = 0, pe_tail = 0, pe_items = 0;
#endif
    for (int w = cluster_id; w < n_items; w += n_clusters) {
      const Item it = decode_item(p, w, cta_rank);
      if (!it.valid) continue;
      const int n_tiles = it.n_tiles;
      uint32_t B[kOwnSlots];  // slot r: columns 64 g + 16 c + r, c = 0..3, of every tile
#pragma unroll
      for (int r = 0; r < kOwnSlots; ++r) B[r] = 0u;
#pragma unroll 1
      for (int t = 0; t < n_tiles; ++t) {
        PROF_T(e0);
        mbar_wait(&bars->tmem_full[as], aphase);
        PROF_T(e1);
        tc_fence_after();
        const uint32_t taddr = tcol + as * kAccCols;
        uint32_t v[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x16(taddr + c * 16, v[c]);
        tmem_wait_ld();
        PROF_T(e2);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars->tmem_empty[as], 0);  // registers hold the tile: stage is free
#pragma unroll
        for (int r = 0; r < kOwnSlots; ++r) {
          const uint32_t m = max(B[r], max(v[0][r], v[1][r]));
          B[r] = max(m, max(v[2][r], v[3][r]));
        }
#ifdef B2M_K1_PROF
        {
          PROF_T(e3);
          PROF_ADD(pe_full, e0, e1);
          PROF_ADD(pe_drain, e1, e2);
          PROF_ADD(pe_rest, e2, e3);
          ++pe_tiles;
        }
#endif
        if (++as == kAccStages) {
          as = 0;
          aphase ^= 1;
        }
      }
      // slot = 16 * group + r; all slot maxima of the item go to the selector warp through shared memory
      // ([slot][row]: conflict-free for both sides); the buffer was consumed a whole item ago
      PROF_T(x0);
      mbar_wait(&bars->sel_empty, sel_phase ^ 1);
#pragma unroll
      for (int r = 0; r < kOwnSlots; ++r) merge[(group * kOwnSlots + r) * kTileM + row_in_cta] = B[r];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->sel_full);
      sel_phase ^= 1;
#ifdef B2M_K1_PROF
      {
        PROF_T(x1);
        PROF_ADD(pe_tail, x0, x1);
        ++pe_items;
      }
#endif
    }
#ifdef B2M_K1_PROF
    if (lane == 0) {
      atomicAdd(&g_k1_prof[4], (unsigned long long)pe_full);
      atomicAdd(&g_k1_prof[5], (unsigned long long)pe_drain);
      atomicAdd(&g_k1_prof[6], (unsigned long long)pe_rest);
      atomicAdd(&g_k1_prof[7], (unsigned long long)pe_tiles);
      atomicAdd(&g_k1_prof[8], (unsigned long long)pe_tail);
      atomicAdd(&g_k1_prof[9], (unsigned long long)pe_items);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs are done with the pair's TMEM / barriers
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kAccStages * kAccCols);
  }
}

// Exact resolution of the candidate rows of one (pair, direction).
// Slot s = 16 g + r holds the columns j = 256 t + 64 g + 16 c + r, c = 0..3, t = 0, 1, ... (slot_col below).
// Candidates are bucketed by winning slot (counting sort in shared memory) so that the n2/64 columns
// of a slot are staged in shared memory ONCE and reused by every candidate of the bucket (a warp per
// candidate, dp4a, oracle scan order per lane, multiset-aware merge across lanes).  Rows whose
// maximum is shared by several slots, and images with more than 16384 features, take the generic
// path that scans global memory.
namespace {

constexpr int kSlotColsMax = 256;      // columns of one slot staged in shared memory (images up to 16384 features)
constexpr int kResolveParts = 4;       // CTAs per (pair, direction); 1/2/4/8 measured within 1.5 % of each other
constexpr int kSlotRowStride = 144;    // bytes; 128-byte descriptors padded so that LDS.128 is conflict-free

// `it`-th column (ascending) of a slot: kSlotCols columns in every 256-column tile
__device__ __forceinline__ int slot_col(int slot, int it) {
  return 256 * (it >> 2) + 64 * (slot >> 4) + 16 * (it & 3) + (slot & 15);
}

__device__ __forceinline__ uint32_t dot128(const uint32_t (&a)[32], const uint4* bp) {
  uint32_t d = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 v = bp[q];
    d = __dp4a(a[4 * q], v.x, d);
    d = __dp4a(a[4 * q + 1], v.y, d);
    d = __dp4a(a[4 * q + 2], v.z, d);
    d = __dp4a(a[4 * q + 3], v.w, d);
  }
  return d;
}

__device__ __forceinline__ void scan_update(uint32_t d, int j, uint32_t& bd, uint32_t& sd, int& bj) {
  if (d > bd) {
    sd = bd;
    bd = d;
    bj = j;
  } else if (d > sd) {
    sd = d;
  }
}

// merge the per-lane scans and take the exact accept decision (lane 0 writes)
__device__ __forceinline__ void finish_candidate(const MatchParams& p, int64_t out_index, uint32_t bd, uint32_t sd,
                                                 int bj, uint32_t best_f, uint32_t s1, bool multi, int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t obd = __shfl_xor_sync(0xffffffffu, bd, o);
    const uint32_t osd = __shfl_xor_sync(0xffffffffu, sd, o);
    const int obj = __shfl_xor_sync(0xffffffffu, bj, o);
    const uint32_t nsd = max(max(sd, osd), min(bd, obd));
    if (obd > bd || (obd == bd && obj >= 0 && (bj < 0 || obj < bj))) {
      bd = obd;
      bj = obj;
    }
    sd = nsd;
  }
  if (lane == 0) {
    const uint32_t second = multi ? sd : max(sd, s1);
    int32_t out = -1;
    if (bd > 0u && bd == best_f) {
      const float fa = __ldg(p.acos_lut + min(bd, 262144u));
      if (!(fa > p.max_distance)) {
        const float fb = __ldg(p.acos_lut + min(second, 262144u));
        if (!(fa >= __fmul_rn(p.max_ratio, fb))) out = bj;
      }
    }
    p.mbuf[out_index] = out;
  }
}

}  // namespace

// dir_only < 0: grid = 2 x n_pairs x kResolveParts, both directions; 0 / 1: grid = n_pairs x kResolveParts, that
// direction only.  With p.gath_desc the rows of direction 1 are the pair's GATHERED descriptors (of image b) and its
// columns are image a (launch_k1_filter_gather).
__global__ void __launch_bounds__(256) b2m_k1_resolve_kernel(const MatchParams p, const uint8_t* __restrict__ desc,
                                                             const int dir_only) {
  // kResolveParts CTAs share the candidates of one (pair, direction): part k takes the slots == k mod
  // kResolveParts (and the rows == k mod kResolveParts of the unstaged candidates), which cuts the
  // serial chain "stage a slot, score its candidates" of a true-match pair by that factor.
  const int part = blockIdx.x % kResolveParts;
  const int pd = blockIdx.x / kResolveParts;
  const int pair = dir_only < 0 ? pd >> 1 : pd;
  const int dir = dir_only < 0 ? pd & 1 : dir_only;
  const int n_cand = p.cand_cnt[pair * 2 + dir];
  if (n_cand == 0) return;
  const bool gathered = dir == 1 && p.gath_desc != nullptr;
  const int ia = p.pairs[2 * pair + dir];
  const int ib = gathered ? p.pairs[2 * pair] : p.pairs[2 * pair + (dir ^ 1)];
  const int nB = p.img_nfeat[ib];
  const int nB_pad = (nB + kRowPad - 1) / kRowPad * kRowPad;
  const uint8_t* A = gathered ? p.gath_desc + static_cast<int64_t>(pair) * p.mstride * kDim
                              : desc + static_cast<int64_t>(p.img_row0[ia]) * kDim;
  const uint8_t* Bm = desc + static_cast<int64_t>(p.img_row0[ib]) * kDim;
  const int64_t base = (static_cast<int64_t>(pair) * 2 + dir) * p.mstride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = nB_pad / kSlots;
  const bool staged = n_items <= kSlotColsMax;

  __shared__ int s_start[kSlots + 1];
  __shared__ int s_fill[kSlots];
  __shared__ __align__(16) uint8_t s_cols[kSlotColsMax * kSlotRowStride];

  // candidate entry = row | slot << 24 (written by the selector warp); a candidate whose maximum is shared
  // by several slots, or any candidate of an image too large for the staging buffer, is "unstaged"
  auto bucket_of = [&](int entry) {
    const uint2 ax = p.aux[base + (entry & 0xFFFFFF)];
    return (ax.x == ax.y || !staged) ? kSlots : (entry >> 24);
  };

  // ---- counting sort of this part's staged candidates by slot
  for (int b = threadIdx.x; b < kSlots; b += 256) s_fill[b] = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < n_cand; c += 256) {
    const int bucket = bucket_of(p.cand_rows[base + c]);
    if (bucket < kSlots) atomicAdd(&s_fill[bucket], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < kSlots; ++b) {  // same offsets in every part: the parts write disjoint ranges
      s_start[b] = acc;
      acc += s_fill[b];
      s_fill[b] = 0;
    }
    s_start[kSlots] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_cand; c += 256) {
    const int entry = p.cand_rows[base + c];
    const int bucket = bucket_of(entry);
    if (bucket < kSlots && bucket % kResolveParts == part)
      p.cand_sorted[base + s_start[bucket] + atomicAdd(&s_fill[bucket], 1)] = entry & 0xFFFFFF;
  }
  __syncthreads();

  // ---- staged slots of this part
  for (int b = part; b < kSlots; b += kResolveParts) {
    const int c0 = s_start[b], c1 = s_start[b + 1];
    if (c0 == c1) continue;  // uniform
    __syncthreads();  // previous slot's readers are done with s_cols
    for (int q = threadIdx.x; q < n_items * 8; q += 256) {
      const int it = q >> 3, seg = q & 7;
      const int j = slot_col(b, it);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(Bm + static_cast<int64_t>(j) * kDim) + seg);
      *reinterpret_cast<uint4*>(s_cols + it * kSlotRowStride + seg * 16) = v;
    }
    __syncthreads();
    for (int c = c0 + warp; c < c1; c += 8) {
      const int row = p.cand_sorted[base + c];
      const uint2 ax = p.aux[base + row];
      uint32_t a[32];
      const uint4* ap = reinterpret_cast<const uint4*>(A + static_cast<int64_t>(row) * kDim);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = __ldg(ap + q);
        a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
      }
      uint32_t bd = 0, sd = 0;
      int bj = -1;
      for (int it = lane; it < n_items; it += 32) {
        const int j = slot_col(b, it);
        const uint32_t d = dot128(a, reinterpret_cast<const uint4*>(s_cols + it * kSlotRowStride));
        scan_update(d, j, bd, sd, bj);
      }
      finish_candidate(p, base + row, bd, sd, bj, ax.x, ax.y, false, lane);
    }
  }

  // ---- unstaged candidates of this part: scan global memory (whole row if several slots share the maximum)
  for (int c = warp; c < n_cand; c += 8) {
    const int entry = p.cand_rows[base + c];
    const int row = entry & 0xFFFFFF;
    if (row % kResolveParts != part) continue;  // uniform in the warp
    const uint2 ax = p.aux[base + row];
    const bool multi = (ax.x == ax.y);
    if (!multi && staged) continue;
    const int slot = entry >> 24;
    uint32_t a[32];
    const uint4* ap = reinterpret_cast<const uint4*>(A + static_cast<int64_t>(row) * kDim);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 v = __ldg(ap + q);
      a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
    uint32_t bd = 0, sd = 0;
    int bj = -1;
    const int n_scan = multi ? nB_pad : n_items;
    for (int it = lane; it < n_scan; it += 32) {
      const int j = multi ? it : slot_col(slot, it);
      const uint32_t d = dot128(a, reinterpret_cast<const uint4*>(Bm + static_cast<int64_t>(j) * kDim));
      scan_update(d, j, bd, sd, bj);
    }
    finish_candidate(p, base + row, bd, sd, bj, ax.x, ax.y, multi, lane);
  }
}

cudaError_t launch_k1_filter(const CUtensorMap& tmap, const MatchParams& p_in,
                             const uint8_t* desc, int n_pairs, int max_strips, int n_dirs, int num_sms,
                             cudaStream_t stream, cudaEvent_t after_filter) {
  // function attributes are per device: several contexts on different GPUs may live in one process
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(b2m_k1_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  MatchParams p = p_in;
  cudaError_t e = cudaMemsetAsync(p.cand_cnt, 0, sizeof(int32_t) * 2 * n_pairs, stream);
  if (e != cudaSuccess) return e;
  // rows never visited by a work item (dir 1 without cross-check is simply not produced)
  p.n_dirs = n_dirs;
  p.blocks_per_image = (max_strips * kTileM + kRowsPerItem - 1) / kRowsPerItem;
  p.n_items = n_pairs * n_dirs * p.blocks_per_image;
  const int clusters = p.n_items < num_sms / kCluster ? p.n_items : num_sms / kCluster;
  if (clusters > 0) {
    b2m_k1_filter_kernel<<<clusters * kCluster, kThreads, kSmemBytes, stream>>>(tmap, tmap, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
#ifdef B2M_K1_PROF
  {
    static int n_launch = 0;
    if (++n_launch % 8 == 0) {  // cumulative counters, printed every 8th launch
      cudaStreamSynchronize(stream);
      unsigned long long h[16];
      cudaMemcpyFromSymbol(h, g_k1_prof, sizeof(h));
      const double mt = h[3] ? double(h[3]) : 1.0, et = h[7] ? double(h[7]) : 1.0, ei = h[9] ? double(h[9]) : 1.0;
      fprintf(stderr,
              "[k1prof] per MMA tile: wait_empty %.0f wait_fullb %.0f issue %.0f | per epi warp-tile: wait_full %.0f "
              "drain %.0f rest %.0f | per item tail %.0f (tiles/item %.1f)\n",
              h[0] / mt, h[1] / mt, h[2] / mt, h[4] / et, h[5] / et, h[6] / et, h[8] / ei, et / ei);
    }
  }
#endif
  if (after_filter) {
    e = cudaEventRecord(after_filter, stream);  // the roofline times the GEMM kernel alone
    if (e != cudaSuccess) return e;
  }
  b2m_k1_resolve_kernel<<<2 * n_pairs * kResolveParts, 256, 0, stream>>>(p, desc, -1);
  return cudaGetLastError();
}

namespace {
// live pair -> (b, a): the column direction becomes the row direction of the swapped pair; dead pair -> dummy
__global__ void b2m_k1_select_dir1_kernel(const int32_t* __restrict__ pairs, const int32_t* __restrict__ cand_cnt,
                                          int n_pairs, int dummy, int32_t* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  const bool live = cand_cnt[2 * k] > 0;
  out[2 * k] = live ? pairs[2 * k + 1] : dummy;
  out[2 * k + 1] = live ? pairs[2 * k] : dummy;
}

__global__ void b2m_compare_matches_kernel(const uint2* __restrict__ arena_a, const int64_t* __restrict__ off_a,
                                           const int32_t* __restrict__ cnt_a, const uint2* __restrict__ arena_b,
                                           const int64_t* __restrict__ off_b, const int32_t* __restrict__ cnt_b,
                                           int32_t* mismatch) {
  const int pair = blockIdx.x;
  const int n = cnt_a[pair];
  if (n != cnt_b[pair]) {
    if (threadIdx.x == 0) atomicOr(mismatch, 1);
    return;
  }
  const uint2* a = arena_a + off_a[pair];
  const uint2* b = arena_b + off_b[pair];
  bool bad = false;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bad = bad || a[i].x != b[i].x || a[i].y != b[i].y;
  if (bad) atomicOr(mismatch, 2);
}
}  // namespace

cudaError_t launch_k1_filter_skip(const CUtensorMap& tmap, const MatchParams& p_in, const uint8_t* desc, int n_pairs,
                                  int max_strips, int num_sms, int32_t* pairs_scratch, int dummy_image,
                                  cudaStream_t stream, cudaEvent_t after_filter) {
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(b2m_k1_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  if (n_pairs <= 0) return cudaSuccess;
  MatchParams p = p_in;
  cudaError_t e = cudaMemsetAsync(p.cand_cnt, 0, sizeof(int32_t) * 2 * n_pairs, stream);
  if (e != cudaSuccess) return e;
  p.n_dirs = 1;
  p.blocks_per_image = (max_strips * kTileM + kRowsPerItem - 1) / kRowsPerItem;
  p.n_items = n_pairs * p.blocks_per_image;
  const int clusters = p.n_items < num_sms / kCluster ? p.n_items : num_sms / kCluster;
  if (clusters > 0) {
    // 1. row direction of every pair
    b2m_k1_filter_kernel<<<clusters * kCluster, kThreads, kSmemBytes, stream>>>(tmap, tmap, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // 2. live pairs swapped, dead pairs -> dummy image (0 features: every work item invalid)
    b2m_k1_select_dir1_kernel<<<(n_pairs + 255) / 256, 256, 0, stream>>>(p.pairs, p.cand_cnt, n_pairs, dummy_image,
                                                                       pairs_scratch);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // 3. column direction of the live pairs = row direction of the swapped pairs, written to the
    //    direction-1 halves: every index in the kernel is (pair * 2 + dir) * mstride (+ row) with dir = 0
    MatchParams q = p;
    q.pairs = pairs_scratch;
    q.mbuf = p.mbuf + p.mstride;
    q.aux = p.aux + p.mstride;
    q.cand_cnt = p.cand_cnt + 1;
    q.cand_rows = p.cand_rows + p.mstride;
    q.cand_sorted = p.cand_sorted + p.mstride;
    b2m_k1_filter_kernel<<<clusters * kCluster, kThreads, kSmemBytes, stream>>>(tmap, tmap, q);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (after_filter) {
    e = cudaEventRecord(after_filter, stream);
    if (e != cudaSuccess) return e;
  }
  // 4. exact resolution of both directions (original pair list and base pointers)
  b2m_k1_resolve_kernel<<<2 * n_pairs * kResolveParts, 256, 0, stream>>>(p, desc, -1);
  return cudaGetLastError();
}

namespace {
// Gathered column direction, step 3 (see match_kernel.cuh): one CTA per pair.  The distinct columns some row of image
// a matched (m12 >= 0 after the exact resolve) are ranked ascending; rank, column list, gathered descriptors (padded
// with zero rows to a multiple of 256) and one work item per 256 gathered rows are written.
constexpr int kGatherMaxWords = 1024;  // columns / 32: images up to 32768 features (SiftMatchingOptions.max_num_matches)
__global__ void __launch_bounds__(256) b2m_k1_gather_kernel(const MatchParams p, const uint8_t* __restrict__ desc,
                                                            uint8_t* __restrict__ gdesc, int32_t* __restrict__ colrank,
                                                            int32_t* __restrict__ cols, int32_t* __restrict__ gcnt,
                                                            int32_t* __restrict__ items, int32_t* __restrict__ n_items,
                                                            const int32_t* __restrict__ enable) {
  const int pair = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ uint32_t s_bits[kGatherMaxWords];
  __shared__ int s_pref[kGatherMaxWords];
  __shared__ int s_warp[8], s_total, s_item0;
  // no candidate, hence no match: the column direction is never consulted (K1); guided matching: pair not eligible
  if (enable ? enable[pair] < 0 : p.cand_cnt[2 * pair] == 0) {
    if (tid == 0) gcnt[pair] = 0;
    return;
  }
  const int ia = p.pairs[2 * pair], ib = p.pairs[2 * pair + 1];
  const int nA = p.img_nfeat[ia], nB = p.img_nfeat[ib];
  const int n_words = (nB + 31) >> 5;
  const int64_t base = static_cast<int64_t>(pair) * p.mstride;
  const int32_t* m12 = p.mbuf + 2 * base;
  for (int w = tid; w < n_words; w += 256) s_bits[w] = 0u;
  __syncthreads();
  for (int i = tid; i < nA; i += 256) {
    const int j = m12[i];
    if (j >= 0) atomicOr(&s_bits[j >> 5], 1u << (j & 31));
  }
  __syncthreads();
  // exclusive prefix of the per-word popcounts (each thread owns a contiguous run of words)
  const int per = (n_words + 255) / 256;
  const int w0 = min(n_words, tid * per), w1 = min(n_words, w0 + per);
  int local = 0;
  for (int w = w0; w < w1; ++w) local += __popc(s_bits[w]);
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int k = 0; k < 8; ++k) {
      const int v = s_warp[k];
      s_warp[k] = acc;
      acc += v;
    }
    s_total = acc;
    const int n_blocks = (acc + kRowPad - 1) / kRowPad;
    s_item0 = n_blocks > 0 ? atomicAdd(n_items, n_blocks) : 0;
    gcnt[pair] = acc;
  }
  __syncthreads();
  {
    int run = s_warp[warp] + incl - local;
    for (int w = w0; w < w1; ++w) {
      s_pref[w] = run;
      run += __popc(s_bits[w]);
    }
  }
  __syncthreads();
  const int nc = s_total;
  const int nc_pad = (nc + kRowPad - 1) / kRowPad * kRowPad;
  for (int t = tid; t < nc_pad / kRowPad; t += 256) {
    items[2 * (s_item0 + t)] = pair;
    items[2 * (s_item0 + t) + 1] = t;
  }
  for (int w = tid; w < n_words; w += 256) {
    uint32_t bits = s_bits[w];
    int r = s_pref[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = (w << 5) + b;
      colrank[base + j] = r;
      cols[base + r] = j;
      ++r;
    }
  }
  __syncthreads();   // cols[] of this CTA are visible to its own threads after the barrier (same block: global writes + __syncthreads)
  const uint4* src0 = reinterpret_cast<const uint4*>(desc + static_cast<int64_t>(p.img_row0[ib]) * kDim);
  uint4* dst0 = reinterpret_cast<uint4*>(gdesc + base * kDim);
  for (int q = tid; q < nc_pad * 8; q += 256) {
    const int r = q >> 3, seg = q & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < nc) v = __ldg(src0 + static_cast<int64_t>(cols[base + r]) * 8 + seg);
    dst0[static_cast<int64_t>(r) * 8 + seg] = v;
  }
}
}  // namespace

// One phase of the gathered schedule (see match_kernel.cuh).  Phases of a batch must be enqueued in order 0, 1, 2, 3 on
// streams ordered by events; splitting them lets the scheduler run phase 1 (ALU kernels that need little of an SM)
// next to the previous batch's RANSAC kernels instead of in front of them.
//   0: row-direction GEMM over all pairs        1: its exact resolve (m12) + gather of the matched columns
//   2: GEMM over the gathered work list          3: exact resolve of the gathered direction
cudaError_t launch_k1_gather_phase(int phase, const CUtensorMap& tmap, const CUtensorMap& tmap_gath, const MatchParams& p_in,
                                   const uint8_t* desc, int n_pairs, int max_strips, int num_sms, const GatherScratch& g,
                                   cudaStream_t stream) {
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(b2m_k1_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  if (n_pairs <= 0) return cudaSuccess;
  MatchParams p = p_in;
  p.n_dirs = 1;
  p.blocks_per_image = (max_strips * kTileM + kRowsPerItem - 1) / kRowsPerItem;
  p.n_items = n_pairs * p.blocks_per_image;
  p.item_list = nullptr;
  p.gath_desc = nullptr;
  const int max_clusters = num_sms / kCluster;
  const int clusters = p.n_items < max_clusters ? p.n_items : max_clusters;
  if (clusters <= 0) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  switch (phase) {
    case 0:
      e = cudaMemsetAsync(p.cand_cnt, 0, sizeof(int32_t) * 2 * n_pairs, stream);
      if (e != cudaSuccess) return e;
      e = cudaMemsetAsync(g.n_items, 0, sizeof(int32_t), stream);
      if (e != cudaSuccess) return e;
      b2m_k1_filter_kernel<<<clusters * kCluster, kThreads, kSmemBytes, stream>>>(tmap, tmap, p);
      break;
    case 1:
      b2m_k1_resolve_kernel<<<n_pairs * kResolveParts, 256, 0, stream>>>(p, desc, 0);
      b2m_k1_gather_kernel<<<n_pairs, 256, 0, stream>>>(p, desc, g.desc, g.colrank, g.cols, g.cnt, g.items, g.n_items, nullptr);
      break;
    case 2: {
      // rows = gathered descriptors, columns = image a; outputs go to the direction-1 halves (every index in the
      // kernel is (pair * 2 + dir) * mstride + row with dir = 0)
      MatchParams q = p;
      q.mbuf = p.mbuf + p.mstride;
      q.aux = p.aux + p.mstride;
      q.cand_cnt = p.cand_cnt + 1;
      q.cand_rows = p.cand_rows + p.mstride;
      q.cand_sorted = p.cand_sorted + p.mstride;
      q.item_list = g.items;
      q.n_items_ptr = g.n_items;
      q.gath_cnt = g.cnt;
      b2m_k1_filter_kernel<<<max_clusters * kCluster, kThreads, kSmemBytes, stream>>>(tmap, tmap_gath, q);
      break;
    }
    default:
      p.gath_desc = g.desc;   // original base pointers: the kernel adds the direction itself
      b2m_k1_resolve_kernel<<<n_pairs * kResolveParts, 256, 0, stream>>>(p, desc, 1);
      break;
  }
  return cudaGetLastError();
}

cudaError_t launch_gather_matched_columns(const MatchParams& p, const uint8_t* desc, int n_pairs, const GatherScratch& g,
                                          const int32_t* enable, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(g.n_items, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  b2m_k1_gather_kernel<<<n_pairs, 256, 0, stream>>>(p, desc, g.desc, g.colrank, g.cols, g.cnt, g.items, g.n_items, enable);
  return cudaGetLastError();
}

cudaError_t launch_k1_filter_gather(const CUtensorMap& tmap, const CUtensorMap& tmap_gath, const MatchParams& p_in,
                                    const uint8_t* desc, int n_pairs, int max_strips, int num_sms, const GatherScratch& g,
                                    cudaStream_t stream, cudaEvent_t after_filter) {
  for (int phase = 0; phase < 4; ++phase) {
    cudaError_t e = launch_k1_gather_phase(phase, tmap, tmap_gath, p_in, desc, n_pairs, max_strips, num_sms, g, stream);
    if (e != cudaSuccess) return e;
    if (phase == 2 && after_filter) {
      e = cudaEventRecord(after_filter, stream);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

cudaError_t launch_compare_matches(const uint2* arena_a, const int64_t* off_a, const int32_t* cnt_a, const uint2* arena_b,
                                   const int64_t* off_b, const int32_t* cnt_b, int n_pairs, int32_t* mismatch,
                                   cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  b2m_compare_matches_kernel<<<n_pairs, 256, 0, stream>>>(arena_a, off_a, cnt_a, arena_b, off_b, cnt_b, mismatch);
  return cudaGetLastError();
}

}  // namespace b2m
