// microbench.cu -- B200 micro-benchmarks that set the roofline denominators of K1:
//   (1) tcgen05.ld (TMEM -> registers) throughput per SM for 4 / 8 / 16 reading warps,
//   (2) tcgen05.mma.kind::i8 128x256x32 issue rate (int8 tensor peak per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../pycolmap_b200/csrc/ptx.cuh"
using namespace b2m;

template <int X>
__device__ __forceinline__ uint32_t ld_chunks(uint32_t taddr) {
  uint32_t acc = 0;
  uint32_t v[32];
#pragma unroll
  for (int c = 0; c < X; ++c) {
    tmem_ld_32x32(taddr + c * 32, v);
    tmem_wait_ld();
#pragma unroll
    for (int r = 0; r < 32; ++r) acc ^= v[r];
  }
  return acc;
}
// variant with all loads of a round in flight before the wait
__device__ __forceinline__ uint32_t ld4_inflight(uint32_t taddr) {
  uint32_t a[32], b[32], c[32], d[32];
  tmem_ld_32x32(taddr, a);
  tmem_ld_32x32(taddr + 32, b);
  tmem_ld_32x32(taddr + 64, c);
  tmem_ld_32x32(taddr + 96, d);
  tmem_wait_ld();
  uint32_t acc = 0;
#pragma unroll
  for (int r = 0; r < 32; ++r) acc ^= a[r] ^ b[r] ^ c[r] ^ d[r];
  return acc;
}

__global__ void ldtm_bw(int iters, int mode, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tbase, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) acc ^= (mode == 0) ? ld_chunks<4>(taddr) : ld4_inflight(taddr);
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

__global__ void __launch_bounds__(128, 1) mma_rate(int iters, int n_tile, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tbase, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the operands -> async proxy (UMMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 32) {
    const uint64_t ad = make_smem_desc_sw128(smem_u32(smem));
    const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 16384));
    const uint32_t idesc = make_idesc_u8u8_s32(128, n_tile);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_i8_ss(tbase + (i & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, k > 0);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

int main() {
  long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, sizeof(long long) * 148);
  cudaMalloc(&d_sink, sizeof(uint32_t) * 148 * 1024);
  long long h[148];
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"sm_clock_attr_khz\": %d", clk_khz);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      ldtm_bw<<<148, warps * 32, 0>>>(iters, mode, d_cycles, d_sink);
      ldtm_bw<<<148, warps * 32, 0>>>(iters, mode, d_cycles, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      const double bytes = double(iters) * warps * 4 * 4096.0;
      printf(", \"ldtm_%s_%dw_bytes_per_clk_per_sm\": %.1f", mode ? "inflight4" : "serial", warps, bytes / avg);
    }
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  for (int n_tile : {256, 128}) {
    mma_rate<<<148, 128, 60000>>>(4000, n_tile, d_cycles);
    mma_rate<<<148, 128, 60000>>>(4000, n_tile, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double macs = 4000.0 * 128 * n_tile * 128;
    printf(", \"i8_mma_n%d_macs_per_clk_per_sm\": %.1f", n_tile, macs / avg);
    // wall-clock rate over the whole chip
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    mma_rate<<<148, 128, 60000>>>(40000, n_tile, d_cycles);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf(", \"i8_mma_n%d_chip_TOPS\": %.1f", n_tile, 2.0 * 40000.0 * 128 * n_tile * 128 * 148 / (ms * 1e-3) / 1e12);
  }
  {
    // SUSTAINED rate: the same MMA loop back to back for ~4 s (the board settles at its power-capped clock), the last
    // second timed.  This is the denominator for a kernel timed inside a long step (B200_PROFILING.md); the burst
    // figures above are for a kernel timed alone.
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int n_tile = 256;
    const double ops_per_launch = 2.0 * 40000.0 * 128 * n_tile * 128 * 148;
    for (int i = 0; i < 280; ++i) mma_rate<<<148, 128, 60000>>>(40000, n_tile, d_cycles);   // ~3 s of warm-up
    cudaEventRecord(a);
    const int timed = 92;                                                                    // ~1 s
    for (int i = 0; i < timed; ++i) mma_rate<<<148, 128, 60000>>>(40000, n_tile, d_cycles);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf(", \"i8_mma_n256_chip_TOPS_sustained\": %.1f, \"sustained_window_ms\": %.0f", ops_per_launch * timed / (ms * 1e-3) / 1e12, ms);
  }
  printf("}\n");
  return 0;
}
