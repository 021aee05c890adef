// microbench3.cu -- what slows the K1 MMA issue loop? N=256 kind::i8 tiles (4 k-steps) with
//   v0: constant descriptors, no commits          v1: B descriptor cycling over 5 stages
//   v2: v1 + two tcgen05.commit per tile           v3: v2 + tcgen05.fence::after_thread_sync per tile
//   v4: v2 + a waiter thread consuming the barriers (mbarrier try_wait traffic)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../pycolmap_b200/csrc/ptx.cuh"
using namespace b2m;

__global__ void __launch_bounds__(384, 1) k(int iters, int variant, long long* cycles, long long* issue_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar_done, bar_a[8], bar_b[8];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 5 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1);
    for (int s = 0; s < 8; ++s) { mbar_init(&bar_a[s], 1); mbar_init(&bar_b[s], 1); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
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
      if (variant >= 2) {
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
  } else if (threadIdx.x == 64 && variant == 4) {
    for (int i = 0; i < iters; ++i) mbar_wait(&bar_a[i & 7], (i >> 3) & 1);
  } else if (warp >= 4 && variant >= 5) {
    // v5: eight warps read TMEM like the K1 epilogue (4 x LDTM.x32 per tile and warp), unsynchronised
    // v6: the same, plus the 64 max3 per tile of the filter epilogue
    const int q = warp & 3, hf = (warp >> 2) & 1;
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
    if (acc == 0x12345) issue_cycles[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
  long long *d_c, *d_i;
  cudaMalloc(&d_c, sizeof(long long) * 148);
  cudaMalloc(&d_i, sizeof(long long) * 148);
  long long h[148], hi[148];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 190000);
  printf("{");
  for (int v = 0; v <= 6; ++v) {
    for (int rep = 0; rep < 2; ++rep) k<<<148, 384, 190000>>>(4000, v, d_c, d_i);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("\"error_v%d\": \"%s\"}\n", v, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_c, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hi, d_i, sizeof(hi), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[i]; b += hi[i]; }
    a /= 148; b /= 148;
    printf("%s\"v%d_cycles_per_tile\": %.1f, \"v%d_issue_cycles_per_tile\": %.1f", v ? ", " : "", v, a / 4000, v, b / 4000);
  }
  printf("}\n");
  return 0;
}
