// five_point_warp.cuh -- the 5-point essential-matrix solver of geom.h (five_point_from_nullspace: Nister's
// elimination, U:estimators/essential_matrix.cc EssentialMatrixFivePointEstimator) as ONE WARP
// per hypothesis instead of one thread per hypothesis.
//
// Why: the serial solver keeps ~6.6 KB of matrices per thread (10 x 20 elimination matrix, the quadratic
// tables of E E^T, the derivative ladder of the degree-10 polynomial).  With 128 hypotheses per CTA and 3 CTAs per
// SM that is 2.5 MB of thread-private state behind a 100-odd KB L1: every multiply-add of the elimination went to
// L2 (B2M_PROF: `solve` 3 M cycles per round, `lo_solve` 160 k cycles per single-thread call).  Here the state of
// one hypothesis lives in the registers of a warp (lane j owns column j of the 10 x 20 matrix) plus 2.6 KB of
// shared memory, the elimination runs on shuffles, and the root finder refines all sign-changing intervals of a
// derivative level at once (one lane per interval).
//
// Same algorithm, same pivoting, same root bracketing as geom.h; only the order of a few floating-point sums
// differs (the rows of the constraint matrix are assembled per monomial instead of per term), so the models agree
// with the serial solver to rounding (tests/test_zz_native_gpu.py::test_warp_five_point_equals_serial_solver).
#pragma once
#include <cstdint>

#include "geom.h"

namespace b2m {
namespace fpw {

// per-warp scratch in shared memory (2.7 KB)
struct Scratch {
  double N[36];          // null-space basis [4][9]: E = x N0 + y N1 + z N2 + N3
  double Q[9][10];       // quadratic tables: cofactors C0..C2 of det(E), then Lambda(0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
  double Mr[6][10];      // rows 4..9, columns 10..19 of the eliminated matrix
  double ladder[10][11]; // derivative ladder of the degree-10 polynomial
  double prev[12], cur[12];
};

// cubic monomials of the 20 columns in Nister's order (gather20 of geom.h) as sorted variable triples,
// variables 0 = x, 1 = y, 2 = z, 3 = 1 (homogenising variable)
static __device__ __constant__ int8_t kTriple[20][3] = {
    {0, 0, 0}, {1, 1, 1}, {0, 0, 1}, {0, 1, 1}, {0, 0, 2}, {0, 0, 3}, {1, 1, 2}, {1, 1, 3}, {0, 1, 2}, {0, 1, 3},
    {0, 2, 2}, {0, 2, 3}, {0, 3, 3}, {1, 2, 2}, {1, 2, 3}, {1, 3, 3}, {2, 2, 2}, {2, 2, 3}, {2, 3, 3}, {3, 3, 3}};
// index of the degree-2 monomial u v (u <= v) in a quadratic table
static __device__ __constant__ int8_t kPair[4][4] = {{0, 3, 4, 6}, {3, 1, 5, 7}, {4, 5, 2, 8}, {6, 7, 8, 9}};
static __device__ __constant__ int8_t kPairU[10] = {0, 1, 2, 0, 0, 1, 0, 1, 2, 3};
static __device__ __constant__ int8_t kPairV[10] = {0, 1, 2, 1, 2, 2, 3, 3, 3, 3};

__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// coefficient of the monomial u v in the product of two linear forms (coefficients of x, y, z, 1)
__device__ __forceinline__ double prod2(const double* N, int ea, int eb, int u, int v) {
  // linear form of entry e: L[e][k] = N[k * 9 + e]
  const double au = N[u * 9 + ea], av = N[v * 9 + ea], bu = N[u * 9 + eb], bv = N[v * 9 + eb];
  return u == v ? au * bu : au * bv + av * bu;
}

__device__ __forceinline__ double eval_sh(const double* q, int deg, double x) {
  double r = q[deg];
  for (int i = deg - 1; i >= 0; --i) r = r * x + q[i];
  return r;
}

// poly_refine of geom.h on a shared-memory polynomial
__device__ __forceinline__ double refine_sh(const double* c, int deg, double lo, double hi, double flo) {
  double x = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    double f = c[deg], df = 0.0;
    for (int i = deg - 1; i >= 0; --i) {
      df = df * x + f;
      f = f * x + c[i];
    }
    if (f == 0.0) return x;
    if ((f < 0.0) == (flo < 0.0)) lo = x; else hi = x;
    double xn = x - f / df;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (fabs(xn - x) <= 4e-16 * fabs(xn) || hi - lo <= 4e-16 * fabs(lo + hi)) return xn;
    x = xn;
  }
  return x;
}

// Real roots of S.ladder[0] (degree <= 10, ascending coefficients) into S.prev (ascending); returns their number.
// poly_real_roots of geom.h with the intervals of a derivative level refined by one lane each.
__device__ inline int real_roots_warp(Scratch& S, int lane) {
  double* c = S.ladder[0];
  double mx = 0.0;
  for (int i = 0; i <= 10; ++i) mx = fmax(mx, fabs(c[i]));
  if (!(mx > 0.0) || !(mx < 1e300)) return 0;
  int deg = 10;
  while (deg > 0 && fabs(c[deg]) <= 1e-14 * mx) --deg;
  if (deg == 0) return 0;
  double bound = 0.0;   // Fujiwara's bound, as geom.h (poly_real_roots): one term per lane
  {
    double t = 0.0;
    if (lane < deg) {
      const double v = fabs(c[lane] / c[deg]);
      if (v > 0.0) t = pow(v, 1.0 / static_cast<double>(deg - lane));
    }
    for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    bound = 2.0 * t * (1.0 + 1e-9);
    if (!(bound > 0.0)) bound = 1.0;
  }
  // derivative ladder: d[k][i] = c[i + k] (i + k)(i + k - 1)...(i + 1), factors applied in that order (as geom.h does)
  for (int e = lane; e < 110; e += 32) {
    const int k = e / 11, i = e % 11;
    if (k >= 1 && k < deg && i <= deg - k) {
      double v = c[i + k];
      for (int t = i + k; t > i; --t) v *= static_cast<double>(t);
      S.ladder[k][i] = v;
    }
  }
  __syncwarp();
  if (lane == 0) S.prev[0] = -S.ladder[deg - 1][0] / S.ladder[deg - 1][1];
  int nprev = 1;
  __syncwarp();
  for (int k = deg - 2; k >= 0; --k) {
    const int dg = deg - k;
    const double* q = S.ladder[k];
    // the intervals the serial scan visits (an interval with hi <= lo is skipped and does not move lo)
    double lo = -bound, my_lo = 0.0, my_hi = 0.0;
    int n_int = 0;
    for (int i = 0; i <= nprev; ++i) {
      const double hi = (i < nprev) ? S.prev[i] : bound;
      if (!(hi > lo)) continue;
      if (n_int == lane) {
        my_lo = lo;
        my_hi = hi;
      }
      ++n_int;
      lo = hi;
    }
    bool emit = false;
    double root = 0.0;
    if (lane < n_int) {
      const double flo = eval_sh(q, dg, my_lo), fhi = eval_sh(q, dg, my_hi);
      if (flo == 0.0) {
        emit = true;
        root = my_lo;
      } else if (fhi != 0.0 && ((flo < 0.0) != (fhi < 0.0))) {
        emit = true;
        root = refine_sh(q, dg, my_lo, my_hi, flo);
      } else if (lane == n_int - 1 && fhi == 0.0) {
        emit = true;
        root = my_hi;
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, emit);
    const int pos = __popc(ballot & ((1u << lane) - 1u));
    const int ncur = min(__popc(ballot), dg);
    __syncwarp();
    if (emit && pos < dg) S.prev[pos] = root;   // every lane has read S.prev (the scan above) before this point
    nprev = ncur;
    __syncwarp();
  }
  return nprev;
}

__device__ __forceinline__ void load_B(const Scratch& S, double (&Bx)[3][4], double (&By)[3][4], double (&Bc)[3][5]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double* a = S.Mr[2 * r];
    const double* b = S.Mr[2 * r + 1];
    Bx[r][0] = a[2];         Bx[r][1] = a[1] - b[2]; Bx[r][2] = a[0] - b[1]; Bx[r][3] = -b[0];
    By[r][0] = a[5];         By[r][1] = a[4] - b[5]; By[r][2] = a[3] - b[4]; By[r][3] = -b[3];
    Bc[r][0] = a[9];         Bc[r][1] = a[8] - b[9]; Bc[r][2] = a[7] - b[8]; Bc[r][3] = a[6] - b[7];
    Bc[r][4] = -b[6];
  }
}

// One warp, first half: the 10 x 20 constraint matrix of the null space S.N, its Gauss-Jordan elimination, and the
// degree-10 polynomial det B(z).  Leaves rows 4..9 / columns 10..19 of the eliminated matrix in S.Mr and the polynomial
// in S.ladder[0]; false when the elimination met a vanishing pivot (no model).  All 32 lanes must call it.
__device__ inline bool eliminate_warp(Scratch& S, int lane) {
  const double* N = S.N;
  // ---- quadratic tables (90 coefficients, three per lane)
  for (int e = lane; e < 90; e += 32) {
    const int q = e / 10, p = e % 10;
    const int u = kPairU[p], v = kPairV[p];
    double val;
    if (q == 0) val = prod2(N, 4, 8, u, v) - prod2(N, 5, 7, u, v);        // C0 = E4 E8 - E5 E7
    else if (q == 1) val = prod2(N, 5, 6, u, v) - prod2(N, 3, 8, u, v);   // C1 = E5 E6 - E3 E8
    else if (q == 2) val = prod2(N, 3, 7, u, v) - prod2(N, 4, 6, u, v);   // C2 = E3 E7 - E4 E6
    else {
      const int s = q - 3;                                                // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
      const int i = s < 3 ? 0 : (s < 5 ? 1 : 2), m = s < 3 ? s : (s < 5 ? s - 2 : 2);
      val = prod2(N, 3 * i, 3 * m, u, v) + prod2(N, 3 * i + 1, 3 * m + 1, u, v) + prod2(N, 3 * i + 2, 3 * m + 2, u, v);
    }
    S.Q[q][p] = val;
  }
  __syncwarp();
  if (lane < 10) {  // Lambda = E E^T - trace(E E^T) / 2 I
    const double half_tr = 0.5 * (S.Q[3][lane] + S.Q[6][lane] + S.Q[8][lane]);
    S.Q[3][lane] -= half_tr;
    S.Q[6][lane] -= half_tr;
    S.Q[8][lane] -= half_tr;
  }
  __syncwarp();
  // ---- lane j < 20 assembles column j of the 10 x 20 constraint matrix: coefficient of its cubic monomial in
  //      sum_m Q_m * L_m = sum over the distinct variables v of the monomial of Q_m[monomial / v] * L_m[v]
  double c[10];
#pragma unroll
  for (int r = 0; r < 10; ++r) c[r] = 0.0;
  if (lane < 20) {
    const int t0 = kTriple[lane][0], t1 = kTriple[lane][1], t2 = kTriple[lane][2];
    // the (at most three) distinct variables of the monomial and the quadratic monomial left when one is taken out
    const int var[3] = {t0, t1, t2};
    const int pr[3] = {kPair[t1][t2], kPair[t0][t2], kPair[t0][t1]};
    const bool use[3] = {true, t1 != t0, t2 != t1};
    const int sym_idx[3][3] = {{3, 4, 5}, {4, 6, 7}, {5, 7, 8}};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      if (!use[t]) continue;
      const int v = var[t], p = pr[t];
#pragma unroll
      for (int m = 0; m < 3; ++m) c[0] += S.Q[m][p] * N[v * 9 + m];                       // det(E): C_m * E_m
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int m = 0; m < 3; ++m) c[1 + i * 3 + j] += S.Q[sym_idx[i][m]][p] * N[v * 9 + (m * 3 + j)];
    }
  }
  // ---- Gauss-Jordan on the first 10 columns, partial pivoting; lane `col` owns the pivot column
  bool singular = false;
#pragma unroll
  for (int col = 0; col < 10; ++col) {
    int piv = col;
    double best = fabs(c[col]);
#pragma unroll
    for (int r = col + 1; r < 10; ++r)
      if (fabs(c[r]) > best) {
        best = fabs(c[r]);
        piv = r;
      }
    piv = __shfl_sync(0xffffffffu, piv, col);
    best = bcast(best, col);
    if (!(best > 1e-300)) {
      singular = true;
      break;
    }
    // row swap col <-> piv in every column
    {
      double pv = c[col];
#pragma unroll
      for (int r = col + 1; r < 10; ++r)
        if (r == piv) {
          const double tmp = c[r];
          c[r] = pv;
          pv = tmp;
        }
      c[col] = pv;
    }
    const double inv = 1.0 / bcast(c[col], col);
    double f[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) f[r] = bcast(c[r], col);   // column `col` after the swap, before the scaling
    c[col] *= inv;
#pragma unroll
    for (int r = 0; r < 10; ++r)
      if (r != col) c[r] -= f[r] * c[col];
  }
  if (singular) return false;
  if (lane >= 10 && lane < 20)
#pragma unroll
    for (int r = 4; r < 10; ++r) S.Mr[r - 4][lane - 10] = c[r];
  __syncwarp();
  // ---- B(z): rows k = e - z f, l = g - z h, m = i - z j; det B(z) = degree-10 polynomial (every lane, redundantly)
  double Bx[3][4], By[3][4], Bc[3][5];
  load_B(S, Bx, By, Bc);
  {
    using geom::detail::pmul;
    double n10[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) n10[i] = 0.0;
    double t7a[8], t7b[8], t6a[7], t6b[7], prod[11];
    pmul(By[1], 3, Bc[2], 4, t7a);
    pmul(Bc[1], 4, By[2], 3, t7b);
#pragma unroll
    for (int i = 0; i < 8; ++i) t7a[i] -= t7b[i];
    pmul(Bx[0], 3, t7a, 7, prod);
#pragma unroll
    for (int i = 0; i < 11; ++i) n10[i] += prod[i];
    pmul(Bx[1], 3, Bc[2], 4, t7a);
    pmul(Bc[1], 4, Bx[2], 3, t7b);
#pragma unroll
    for (int i = 0; i < 8; ++i) t7a[i] -= t7b[i];
    pmul(By[0], 3, t7a, 7, prod);
#pragma unroll
    for (int i = 0; i < 11; ++i) n10[i] -= prod[i];
    pmul(Bx[1], 3, By[2], 3, t6a);
    pmul(By[1], 3, Bx[2], 3, t6b);
#pragma unroll
    for (int i = 0; i < 7; ++i) t6a[i] -= t6b[i];
    pmul(Bc[0], 4, t6a, 6, prod);
#pragma unroll
    for (int i = 0; i < 11; ++i) n10[i] += prod[i];
    if (lane == 0)
#pragma unroll
      for (int i = 0; i < 11; ++i) S.ladder[0][i] = n10[i];
  }
  __syncwarp();
  return true;
}

// The (x, y) of one root z of det B(z) and its essential matrix E = x N0 + y N1 + z N2 + N3; false when B(z) has no
// usable null vector.  Bx / By / Bc as load_B gives them.
__device__ __forceinline__ bool model_of_root(const double (&Bx)[3][4], const double (&By)[3][4], const double (&Bc)[3][5],
                                              const double* N, double z, double* E) {
  double bx[3], by[3], bc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    bx[k] = geom::poly_eval(Bx[k], 3, z);
    by[k] = geom::poly_eval(By[k], 3, z);
    bc[k] = geom::poly_eval(Bc[k], 4, z);
  }
  double bestw = 0.0, X = 0.0, Y = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int b = (a + 1) % 3;
    const double cxp = by[a] * bc[b] - bc[a] * by[b];
    const double cyp = bc[a] * bx[b] - bx[a] * bc[b];
    const double cw = bx[a] * by[b] - by[a] * bx[b];
    if (fabs(cw) > fabs(bestw)) {
      bestw = cw;
      X = cxp;
      Y = cyp;
    }
  }
  if (!(fabs(bestw) > 0.0)) return false;
  const double x = X / bestw, y = Y / bestw;
#pragma unroll
  for (int e = 0; e < 9; ++e) E[e] = x * N[e] + y * N[9 + e] + z * N[18 + e] + N[27 + e];
  return true;
}

// One THREAD, second half (minimal samples: 100-odd independent hypotheses per CTA, where the latency-bound root
// refinement wants thread-level parallelism): real roots of the polynomial (serial bracketing of geom.h) and one model
// per root.  N [4][9], poly [11], Mr [6][10] are what eliminate_warp left (any memory space); models [<= 10][9].
__device__ inline int finish_thread(const double* N, const double* poly, const double* Mr, double* models) {
  double roots[10];
  const int nr = geom::poly_real_roots(poly, 10, roots);
  if (nr == 0) return 0;
  double Bx[3][4], By[3][4], Bc[3][5];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double* a = Mr + 20 * r;
    const double* b = Mr + 20 * r + 10;
    Bx[r][0] = a[2];         Bx[r][1] = a[1] - b[2]; Bx[r][2] = a[0] - b[1]; Bx[r][3] = -b[0];
    By[r][0] = a[5];         By[r][1] = a[4] - b[5]; By[r][2] = a[3] - b[4]; By[r][3] = -b[3];
    Bc[r][0] = a[9];         Bc[r][1] = a[8] - b[9]; Bc[r][2] = a[7] - b[8]; Bc[r][3] = a[6] - b[7];
    Bc[r][4] = -b[6];
  }
  int nm = 0;
  for (int r = 0; r < nr; ++r)
    if (model_of_root(Bx, By, Bc, N, roots[r], models + 9 * nm)) ++nm;
  return nm;
}

// One warp: the essential matrices of the null space S.N.  models: [<= 10][9] (shared or global); returns the
// number of models (the same value in every lane).  All 32 lanes must call it.  (Local optimisation: one solve at a
// time, so the root refinement runs one lane per interval.)
__device__ inline int five_point_warp(Scratch& S, double* models, int lane) {
  if (!eliminate_warp(S, lane)) return 0;
  const double* N = S.N;
  double Bx[3][4], By[3][4], Bc[3][5];
  const int nr = real_roots_warp(S, lane);
  load_B(S, Bx, By, Bc);
  // ---- one lane per root: (x, y) from the null vector of B(z), E = x N0 + y N1 + z N2 + N3
  bool ok = false;
  double E[9];
  if (lane < nr) ok = model_of_root(Bx, By, Bc, N, S.prev[lane], E);
  const unsigned ballot = __ballot_sync(0xffffffffu, ok);
  if (ok) {
    double* out = models + 9 * __popc(ballot & ((1u << lane) - 1u));
#pragma unroll
    for (int e = 0; e < 9; ++e) out[e] = E[e];
  }
  __syncwarp();
  return __popc(ballot);
}

}  // namespace fpw
}  // namespace b2m
