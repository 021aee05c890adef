// microbench2.cu -- does shared-memory bandwidth cap SS-mode kind::i8 MMAs at K=128?
//   modes: 0 = SS (A, B from smem), 1 = TS (A from TMEM); optional background smem writers emulate
//   the TMA fill traffic (bytes per MMA-tile written by 4 extra warps).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../pycolmap_b200/csrc/ptx.cuh"
using namespace b2m;

__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(256, 1) mma_bw(int iters, int n_tile, int mode, int writers, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (32768 + 8 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_u8u8_s32(128, n_tile);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // two strips x 4 k-steps against B stage (i % 8), like K1
      const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 32768 + (i & 7) * 16384));
      for (int s = 0; s < 2; ++s) {
        const uint64_t ad = make_smem_desc_sw128(smem_u32(smem + s * 16384));
        const uint32_t d = tbase + (i & 1) * 192 + s * 96;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 0) mma_i8_ss(d, ad + 2 * k, bd + 2 * k, idesc, k > 0);
          else mma_i8_ts(d, tbase + 384 + s * 32 + 8 * k, bd + 2 * k, idesc, k > 0);
        }
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    done = 1;
  } else if (warp >= 4 && writers) {
    // background writers: 128 threads x 16 B = 2 KiB per round into the B ring
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    uint8_t* base = smem + 32768;
    int r = 0;
    while (!done) {
      *reinterpret_cast<uint4*>(base + ((r * 2048 + (threadIdx.x - 128) * 16) & (8 * 16384 - 1))) = v;
      ++r;
      if (writers > 1) __nanosleep(writers);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 148);
  long long h[148];
  cudaFuncSetAttribute(mma_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, 170000);
  printf("{");
  bool first = true;
  for (int mode = 0; mode < 2; ++mode)
    for (int n_tile : {128, 96, 64})
      for (int writers : {0, 1}) {
        for (int rep = 0; rep < 2; ++rep) mma_bw<<<148, 256, 170000>>>(2000, n_tile, mode, writers, d_cycles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s\"error_mode%d_n%d\": \"%s\"", first ? "" : ", ", mode, n_tile, cudaGetErrorString(e)); printf("}\n"); return 1; }
        cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        const double macs = 2000.0 * 2 * 128 * n_tile * 128;
        printf("%s\"%s_n%d_w%d_macs_per_clk\": %.0f", first ? "" : ", ", mode ? "TS" : "SS", n_tile, writers, macs / avg);
        first = false;
      }
  printf("}\n");
  return 0;
}
