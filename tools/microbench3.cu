// microbench3.cu -- what slows the K1 MMA issue loop? N=256 kind::i8 tiles (4 k-steps) with
//   v0: constant descriptors, no commits          v1: B descriptor cycling over 5 stages
//   v2: v1 + two tcgen05.commit per tile           v3: v2 + tcgen05.fence::after_thread_sync per tile
//   v4: v2 + a waiter thread consuming the barriers (mbarrier try_wait traffic)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../pycolmap_b200/csrc/ptx.cuh"
using namespace b2m;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) k(const __grid_constant__ CUtensorMap tmap, int iters, int variant, long long* cycles, long long* issue_cycles, long long* epi_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar_done, bar_a[8], bar_b[8], bar_t[5];
  __shared__ volatile int mma_done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 5 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    for (int s = 0; s < 8; ++s) { mbar_init(&bar_a[s], variant >= 7 ? 2 : 1); mbar_init(&bar_b[s], 1); }
    for (int s = 0; s < 5; ++s) mbar_init(&bar_t[s], 1);
    mma_done = 0;
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_u8u8_s32(128, 256);
    const uint64_t ad = make_smem_desc_sw128(smem_u32(smem));
    const long long t0 = clock64();
    long long issue = 0;
    for (int i = 0; i < iters; ++i) {
      const int st = (variant >= 1) ? (i % 5) : 0;
      const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 16384 + st * 32768));
      if (variant >= 3) tc_fence_after();
      const long long c0 = clock64();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) mma_i8_ss(tbase + (i & 1) * 256, ad + 2 * kk, bd + 2 * kk, idesc, kk > 0);
      if (variant == 7) {         // v7: multicast commits (cluster of 2, cta_group::1), like K1 v5-v7
        mma_commit_multicast(&bar_a[i & 7], 3);
        mma_commit(&bar_b[i & 7]);
      } else if (variant == 8) {  // v8: ONE multicast commit per tile
        mma_commit_multicast(&bar_a[i & 7], 3);
      } else if (variant >= 2) {
        mma_commit(&bar_a[i & 7]);
        mma_commit(&bar_b[i & 7]);
      }
      issue += clock64() - c0;
    }
    mma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    issue_cycles[blockIdx.x] = issue;
    mma_done = 1;
  } else if (threadIdx.x == 64 && variant == 4) {
    for (int i = 0; i < iters; ++i) mbar_wait(&bar_a[i & 7], (i >> 3) & 1);
  } else if (threadIdx.x == 96 && variant == 9) {
    // v9: a TMA producer streams 32 KiB tiles into the 5-stage B ring while the MMAs run (no data dependency)
    uint32_t st = 0, ph = 0;
    long long n = 0;
    while (!mma_done) {
      mbar_arrive_expect_tx(&bar_t[st], 32768);
      tma_load_2d(smem + 16384 + st * 32768, &tmap, &bar_t[st], 0, static_cast<int>((n * 256) % 65536));
      tma_load_2d(smem + 16384 + st * 32768 + 16384, &tmap, &bar_t[st], 0, static_cast<int>((n * 256 + 128) % 65536));
      mbar_wait(&bar_t[st], ph);
      if (++st == 5) { st = 0; ph ^= 1; }
      ++n;
    }
    epi_cycles[blockIdx.x] = n;   // tiles streamed
  } else if (warp >= 4 && variant >= 5 && variant <= 6) {
    // v5: eight warps read TMEM like the K1 epilogue (4 x LDTM.x32 per tile and warp), unsynchronised
    // v6: the same, plus the 64 max3 per tile of the filter epilogue
    const int q = warp & 3, hf = (warp >> 2) & 1;
    const long long e0 = clock64();
    uint32_t B0[32], acc = 0;
#pragma unroll
    for (int r = 0; r < 32; ++r) B0[r] = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t taddr = tbase + (static_cast<uint32_t>(q * 32) << 16) + ((i + 1) & 1) * 256 + hf * 128;
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr, va);
      tmem_ld_32x32(taddr + 32, vb);
      tmem_wait_ld();
      if (variant >= 6) {
#pragma unroll
        for (int r = 0; r < 32; ++r) B0[r] = max(B0[r], max(va[r], vb[r]));
      } else acc ^= va[0] ^ vb[5];
      tmem_ld_32x32(taddr + 64, va);
      tmem_ld_32x32(taddr + 96, vb);
      tmem_wait_ld();
      if (variant >= 6) {
#pragma unroll
        for (int r = 0; r < 32; ++r) B0[r] = max(B0[r], max(va[r], vb[r]));
      } else acc ^= va[3] ^ vb[7];
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) acc ^= B0[r];
    if (threadIdx.x == 128) epi_cycles[blockIdx.x] = clock64() - e0;
    if (acc == 0x12345) issue_cycles[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* dbuf; cudaMalloc(&dbuf, 66000ull * 128); cudaMemset(dbuf, 7, 66000ull * 128);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap tmap; cuuint64_t gdim[2] = {128, 66000}; cuuint64_t gstr[1] = {128}; cuuint32_t box[2] = {128, 128}; cuuint32_t es[2] = {1, 1};
  ((PFN_enc)fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dbuf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  long long *d_c, *d_i, *d_e;
  cudaMalloc(&d_e, sizeof(long long) * 148);
  cudaMemset(d_e, 0, sizeof(long long) * 148);
  cudaMalloc(&d_c, sizeof(long long) * 148);
  cudaMalloc(&d_i, sizeof(long long) * 148);
  long long h[148], hi[148];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 190000);
  printf("{");
  for (int v : {0, 6, 9}) {
    for (int rep = 0; rep < 2; ++rep) k<<<148, 384, 190000>>>(tmap, 4000, v, d_c, d_i, d_e);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("\"error_v%d\": \"%s\"}\n", v, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_c, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hi, d_i, sizeof(hi), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[i]; b += hi[i]; }
    a /= 148; b /= 148;
    long long he[148]; cudaMemcpy(he, d_e, sizeof(he), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += he[i]; c /= 148;
    printf("%s\"v%d_cycles_per_tile\": %.1f, \"v%d_issue_cycles_per_tile\": %.1f, \"v%d_epi_cycles_per_tile\": %.1f", v ? ", " : "", v, a / 4000, v, b / 4000, v, c / 4000);
  }
  printf("}\n");
  return 0;
}
