// eig_warp.cuh -- the K smallest eigenvectors of the 9 x 9 normal matrix of an LO refit, computed by ONE WARP.
//
// Same algorithm, same operations in the same order per matrix element as geom.h smallest_eigvecs_invit<K>
// (Cholesky-based inverse subspace iteration), so the result is bit-identical; what changes is who computes what:
// the serial version keeps L[81] in thread-local memory and walks one ~1500-operation dependent chain while the other
// 127 threads of the CTA wait at a barrier (B2M_PROF `lo_solve`: 17 % of the CTA time of the RANSAC kernels).  Here
// the matrix lives in shared memory, lane i owns row i (and column i) of the factor in registers, the column of a
// Cholesky step and the updates of a substitution step run in parallel, and only the inherently serial parts (one
// pivot, one solved component at a time) stay serial.
#pragma once
#include <cstdint>

namespace b2m {
namespace eigw {

struct Scratch {
  double L[81];      // the factor, row-major
  double v[4][9];    // the iterated basis
  double inv;        // 1 / L(j, j) of the current Cholesky step
};

__device__ __forceinline__ double bcast(double x, int src) {
  const int lo = __shfl_sync(0xffffffffu, __double2loint(x), src);
  const int hi = __shfl_sync(0xffffffffu, __double2hiint(x), src);
  return __hiloint2double(hi, lo);
}

// S45: upper triangle of the symmetric matrix, row-major packed (geom.h sym9_expand); out: [K][9], out[0] belongs to the
// smallest eigenvalue.  All 32 lanes of the warp must call; S45 / out may live in shared or global memory.
template <int K>
__device__ __forceinline__ void smallest_eigvecs_warp(const double* S45, double* out, Scratch& W, int lane) {
  // ---- expand + regularise (lane i: row i)
  if (lane < 9) {
    for (int j = 0; j < 9; ++j) {
      const int a = lane < j ? lane : j, b = lane < j ? j : lane;      // packed index of (a, b), a <= b
      W.L[lane * 9 + j] = S45[a * 9 - a * (a - 1) / 2 + (b - a)];
    }
  }
  __syncwarp();
  double tr = 0.0;
  for (int i = 0; i < 9; ++i) tr += W.L[i * 9 + i];
  const double mu = tr * 1e-14 + 1e-300;
  __syncwarp();
  if (lane < 9) W.L[lane * 9 + lane] += mu;
  __syncwarp();
  // ---- in-place Cholesky, lower triangle: the pivot on lane j, the rest of column j on lanes j + 1 .. 8
#pragma unroll 1
  for (int j = 0; j < 9; ++j) {
    if (lane == j) {
      double d = W.L[j * 9 + j];
      for (int k = 0; k < j; ++k) d -= W.L[j * 9 + k] * W.L[j * 9 + k];
      if (!(d > mu * 1e-3)) d = mu * 1e-3;
      const double ljj = sqrt(d);
      W.L[j * 9 + j] = ljj;
      W.inv = 1.0 / ljj;
    }
    __syncwarp();
    if (lane > j && lane < 9) {
      double s = W.L[lane * 9 + j];
      for (int k = 0; k < j; ++k) s -= W.L[lane * 9 + k] * W.L[j * 9 + k];
      W.L[lane * 9 + j] = s * W.inv;
    }
    __syncwarp();
  }
  // ---- row / column of this lane in registers, reciprocal diagonal
  const int li = lane < 9 ? lane : 8;
  double row[9], col[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    row[j] = W.L[li * 9 + j];
    col[j] = W.L[j * 9 + li];
  }
  const double rd = 1.0 / W.L[li * 9 + li];   // 1 / L(i, i)
  // deterministic, generic start vectors
  if (lane < 9)
    for (int k = 0; k < K; ++k) {
      const int h = (lane * 37 + k * 101 + 11) % 17;
      W.v[k][lane] = (static_cast<double>(h) - 8.0) * 0.1 + (lane == 8 - k ? 1.0 : 0.0);
    }
  __syncwarp();
#pragma unroll 1
  for (int it = 0; it < 6; ++it) {
    double before[9];
    if (K == 1)
      for (int i = 0; i < 9; ++i) before[i] = W.v[0][i];
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      // L y = v: component j is final once the components before it are; every later row then takes its update
      // (row i subtracts L(i, j) y_j in increasing j, the serial order)
      double s = W.v[k][li];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const double yj = bcast(s * rd, j);
        if (lane > j) s -= row[j] * yj;
        if (lane == j) s = yj;
      }
      // L^T x = y: x_i = (y_i - sum_{j > i} L(j, i) x_j) / L(i, i) with the sum in increasing j (the serial order), so
      // lane i waits for all later components and then runs its own short chain
      double x[9];
#pragma unroll
      for (int i = 8; i >= 0; --i) {
        double t = s;
#pragma unroll
        for (int j = i + 1; j < 9; ++j) t -= col[j] * x[j];
        x[i] = bcast(t * rd, i);
      }
      __syncwarp();
      if (lane < 9) {
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) mine = (i == lane) ? x[i] : mine;
        W.v[k][lane] = mine;
      }
      __syncwarp();
    }
    // modified Gram-Schmidt + normalisation: 9-term serial sums, computed redundantly by every lane (same bits)
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      for (int q = 0; q < k; ++q) {
        double dot = 0.0;
        for (int i = 0; i < 9; ++i) dot += W.v[q][i] * W.v[k][i];
        __syncwarp();
        if (lane < 9) W.v[k][lane] -= dot * W.v[q][lane];
        __syncwarp();
      }
      double nn = 0.0;
      for (int i = 0; i < 9; ++i) nn += W.v[k][i] * W.v[k][i];
      const double inv = 1.0 / sqrt(nn > 0.0 ? nn : 1.0);
      __syncwarp();
      if (lane < 9) W.v[k][lane] *= inv;
      __syncwarp();
    }
    if (K == 1 && it > 0) {
      double dp = 0.0, dm = 0.0;   // the iterate may flip its sign from step to step
      for (int i = 0; i < 9; ++i) {
        const double a = W.v[0][i];
        dp += (a - before[i]) * (a - before[i]);
        dm += (a + before[i]) * (a + before[i]);
      }
      if ((dp < dm ? dp : dm) < 1e-30) break;   // uniform across the warp: every lane computed the same sums
    }
  }
  if (lane < 9)
    for (int k = 0; k < K; ++k) out[k * 9 + lane] = W.v[k][lane];
  __syncwarp();
}

}  // namespace eigw
}  // namespace b2m
