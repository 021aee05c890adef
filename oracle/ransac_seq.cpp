// ransac_seq.cpp -- TEST INFRASTRUCTURE ONLY (CPU baseline arm of bench.py, cross-checked in tests/).
//
// Scalar fp64, SEQUENTIAL LO-RANSAC two-view verification with the control flow of the reference's CPU path:
// EstimateTwoViewGeometry -> EstimateCalibrated / Uncalibrated TwoViewGeometry (U:estimators/two_view_geometry.cc,
// reached from R:estimators/two_view_geometry.h:95-151 and from the VerifierWorker threads behind
// R:pipeline/match_features.h:45-48), LORANSAC<Estimator, LocalEstimator>::Estimate (U:optim/loransac.h): one
// hypothesis at a time, every model scored over all matches, local optimisation on a new best, dynamic trial bound.
// One image pair per thread (orc_verify_pairs), like upstream's pool of verifier threads.
//
// The minimal / non-minimal solver math (5-point, 7-point, 8-point, DLT, Sampson / transfer residuals) comes from
// pycolmap_b200/csrc/geom.h compiled for the host -- it is plain header-only C++ and is itself checked against the
// numpy restatement oracle/ransac.py (tests/test_oracle_ransac.py).  The RANSAC loop, sampler, decision tree and
// watermark test below are a second, independent restatement of what oracle/ransac.py states in numpy; the two must
// agree statistically (tests/test_oracle_ransac.py::test_sequential_cpp_oracle_agrees_with_numpy_oracle).
// PARITY UNPINNED like the rest of oracle/ (the reference holds no vectors for this path).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "../pycolmap_b200/csrc/geom.h"

using namespace b2m::geom;

namespace {

struct Opts {
  int min_num_inliers;
  double min_E_F_inlier_ratio, max_H_inlier_ratio, watermark_min_inlier_ratio, watermark_border_size;
  int detect_watermark, force_H_use;
  double max_error, min_inlier_ratio, confidence, dyn_mult;
  int min_num_trials, max_num_trials;
};

struct Report {
  bool success = false;
  int num_inliers = 0;
  double residual_sum = 1e300;
  double model[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<char> mask;
  long models_scored = 0;
};

// kind: 0 E (5-point both), 1 F (7-point / 8-point), 2 H (DLT both), 3 2-D translation
struct Problem {
  int kind;
  const std::vector<double>*x1, *y1, *x2, *y2;
  int n() const { return static_cast<int>(x1->size()); }
  int k_min() const { return kind == 0 ? 5 : kind == 1 ? 7 : kind == 2 ? 4 : 1; }
  int k_local() const { return kind == 0 ? 5 : kind == 1 ? 8 : kind == 2 ? 4 : 1; }
  double residual(const double* M, int i) const {
    if (kind == 2) return homography_sq(M, (*x1)[i], (*y1)[i], (*x2)[i], (*y2)[i]);
    if (kind == 3) {
      const double ex = (*x2)[i] - ((*x1)[i] + M[0]), ey = (*y2)[i] - ((*y1)[i] + M[1]);
      return ex * ex + ey * ey;
    }
    return sampson_sq(M, (*x1)[i], (*y1)[i], (*x2)[i], (*y2)[i]);
  }
  // models from the points with the given indices; `local`: the non-minimal estimator
  int estimate(const int* idx, int m, bool local, double* models) const {
    std::vector<double> a(m), b(m), c(m), d(m);
    for (int i = 0; i < m; ++i) {
      a[i] = (*x1)[idx[i]]; b[i] = (*y1)[idx[i]]; c[i] = (*x2)[idx[i]]; d[i] = (*y2)[idx[i]];
    }
    switch (kind) {
      case 0: return estimate_E(a.data(), b.data(), c.data(), d.data(), m, models);
      case 1: return local ? estimate_F8(a.data(), b.data(), c.data(), d.data(), m, models)
                           : estimate_F7(a.data(), b.data(), c.data(), d.data(), models);
      case 2: return estimate_H(a.data(), b.data(), c.data(), d.data(), m, models);
      default: {
        double sx = 0, sy = 0;
        for (int i = 0; i < m; ++i) { sx += c[i] - a[i]; sy += d[i] - b[i]; }
        for (int k = 0; k < 9; ++k) models[k] = 0.0;
        models[0] = sx / m; models[1] = sy / m;
        return 1;
      }
    }
  }
};

bool better(int c, double s, int bc, double bs) { return c > bc || (c == bc && s < bs); }

// LORANSAC<Estimator, LocalEstimator>::Estimate, sequential
Report loransac(const Problem& P, const Opts& o, double max_error, double min_inlier_ratio, std::mt19937& rng) {
  Report rep;
  const int n = P.n();
  rep.mask.assign(n, 0);
  if (n < P.k_min()) return rep;
  const double max_residual = max_error * max_error;
  const double clip = compute_num_trials(std::floor(min_inlier_ratio * 100000.0), 100000.0, o.confidence, o.dyn_mult, P.k_min());
  const int max_trials = static_cast<int>(std::min<double>(o.max_num_trials, clip));
  double dyn_max = max_trials;
  int best_cnt = 0;
  double best_sum = 1e300;
  bool have = false;
  std::vector<int> perm(n), inl;
  for (int i = 0; i < n; ++i) perm[i] = i;
  std::vector<double> res(n);
  double models[90], lmodels[90];
  bool abort = false;
  for (int trial = 0; trial < max_trials && !abort; ++trial) {
    const int k = P.k_min();
    for (int i = 0; i < k; ++i) {  // RandomSampler: partial Fisher-Yates on a persistent index vector
      std::uniform_int_distribution<int> pick(i, n - 1);
      std::swap(perm[i], perm[pick(rng)]);
    }
    const int nm = P.estimate(perm.data(), k, false, models);
    for (int m = 0; m < nm && !abort; ++m) {
      const double* M = models + 9 * m;
      int c = 0;
      double s = 0.0;
      for (int i = 0; i < n; ++i) {
        res[i] = P.residual(M, i);
        if (res[i] <= max_residual) { ++c; s += res[i]; }
      }
      ++rep.models_scored;
      if (better(c, s, best_cnt, best_sum)) {
        best_cnt = c; best_sum = s; have = true;
        std::memcpy(rep.model, M, sizeof(rep.model));
        if (c > P.k_min() && c >= P.k_local()) {
          for (int lt = 0; lt < 10; ++lt) {
            inl.clear();
            for (int i = 0; i < n; ++i)
              if (P.residual(rep.model, i) <= max_residual) inl.push_back(i);
            const int prev = best_cnt;
            const int nl = P.estimate(inl.data(), static_cast<int>(inl.size()), true, lmodels);
            for (int q = 0; q < nl; ++q) {
              int lc = 0;
              double ls = 0.0;
              for (int i = 0; i < n; ++i) {
                const double r = P.residual(lmodels + 9 * q, i);
                if (r <= max_residual) { ++lc; ls += r; }
              }
              ++rep.models_scored;
              if (better(lc, ls, best_cnt, best_sum)) {
                best_cnt = lc; best_sum = ls;
                std::memcpy(rep.model, lmodels + 9 * q, sizeof(rep.model));
              }
            }
            if (best_cnt <= prev) break;
          }
        }
        dyn_max = compute_num_trials(best_cnt, n, o.confidence, o.dyn_mult, P.k_min());
      }
      if (trial >= dyn_max && trial >= o.min_num_trials) abort = true;
    }
  }
  rep.num_inliers = best_cnt;
  rep.residual_sum = best_sum;
  if (!have || best_cnt < P.k_min()) return rep;
  rep.success = true;
  for (int i = 0; i < n; ++i) rep.mask[i] = P.residual(rep.model, i) <= max_residual;
  return rep;
}

struct Cam { double fx, fy, cx, cy; int w, h, prior; };

enum { UNDEFINED = 0, DEGENERATE, CALIBRATED, UNCALIBRATED, PLANAR, PANORAMIC, PLANAR_OR_PANORAMIC, WATERMARK };

struct Result {
  int config = DEGENERATE, nE = 0, nF = 0, nH = 0;
  long models_scored = 0;
  double E[9] = {0}, F[9] = {0}, H[9] = {0};
  std::vector<uint32_t> inliers;
};

Result estimate_tvg(const Cam& c1, const double* p1, const Cam& c2, const double* p2, const uint32_t* matches, int64_t m,
                    const Opts& o, uint32_t seed) {
  Result R;
  if (m < o.min_num_inliers) return R;
  std::mt19937 rng(seed);
  std::vector<double> x1(m), y1(m), x2(m), y2(m), nx1(m), ny1(m), nx2(m), ny2(m);
  for (int64_t i = 0; i < m; ++i) {
    const uint32_t a = matches[2 * i], b = matches[2 * i + 1];
    x1[i] = p1[2 * a]; y1[i] = p1[2 * a + 1]; x2[i] = p2[2 * b]; y2[i] = p2[2 * b + 1];
    nx1[i] = (x1[i] - c1.cx) / c1.fx; ny1[i] = (y1[i] - c1.cy) / c1.fy;   // CamFromImg, pinhole models
    nx2[i] = (x2[i] - c2.cx) / c2.fx; ny2[i] = (y2[i] - c2.cy) / c2.fy;
  }
  const bool calibrated = c1.prior && c2.prior && !o.force_H_use;
  Report rE, rF, rH;
  if (calibrated) {
    const double e = 0.5 * (o.max_error / (0.5 * (c1.fx + c1.fy)) + o.max_error / (0.5 * (c2.fx + c2.fy)));
    rE = loransac(Problem{0, &nx1, &ny1, &nx2, &ny2}, o, e, o.min_inlier_ratio, rng);
    std::memcpy(R.E, rE.model, sizeof(R.E));
  }
  if (!o.force_H_use) {
    rF = loransac(Problem{1, &x1, &y1, &x2, &y2}, o, o.max_error, o.min_inlier_ratio, rng);
    std::memcpy(R.F, rF.model, sizeof(R.F));
  }
  rH = loransac(Problem{2, &x1, &y1, &x2, &y2}, o, o.max_error, o.min_inlier_ratio, rng);
  std::memcpy(R.H, rH.model, sizeof(R.H));
  R.nE = rE.num_inliers; R.nF = rF.num_inliers; R.nH = rH.num_inliers;
  R.models_scored = rE.models_scored + rF.models_scored + rH.models_scored;
  const int mn = o.min_num_inliers, nE = R.nE, nF = R.nF, nH = R.nH;
  const std::vector<char>* mask = nullptr;
  int num = 0;
  if (o.force_H_use) {
    if (!rH.success || nH < mn) return R;
    R.config = PLANAR_OR_PANORAMIC; mask = &rH.mask; num = nH;
  } else if (!calibrated) {  // EstimateUncalibratedTwoViewGeometry: F's mask always
    if ((!rF.success && !rH.success) || (nF < mn && nH < mn)) return R;
    R.config = (static_cast<double>(nH) / nF > o.max_H_inlier_ratio) ? PLANAR_OR_PANORAMIC : UNCALIBRATED;
    mask = &rF.mask; num = nF;
  } else {
    if ((!rE.success && !rF.success && !rH.success) || (nE < mn && nF < mn && nH < mn)) return R;
    const double E_F = static_cast<double>(nE) / nF, H_F = static_cast<double>(nH) / nF, H_E = static_cast<double>(nH) / nE;
    if (rE.success && E_F > o.min_E_F_inlier_ratio && nE >= mn) {
      if (nE >= nF) { mask = &rE.mask; num = nE; } else { mask = &rF.mask; num = nF; }
      if (H_E > o.max_H_inlier_ratio) {
        R.config = PLANAR_OR_PANORAMIC;
        if (nH > num) { mask = &rH.mask; num = nH; }
      } else {
        R.config = CALIBRATED;
      }
    } else if (rF.success && nF >= mn) {
      mask = &rF.mask; num = nF;
      if (H_F > o.max_H_inlier_ratio) {
        R.config = PLANAR_OR_PANORAMIC;
        if (nH > num) { mask = &rH.mask; num = nH; }
      } else {
        R.config = UNCALIBRATED;
      }
    } else if (rH.success && nH >= mn) {
      mask = &rH.mask; num = nH; R.config = PLANAR_OR_PANORAMIC;
    } else {
      return R;
    }
  }
  for (int64_t i = 0; i < m; ++i)
    if ((*mask)[i]) { R.inliers.push_back(matches[2 * i]); R.inliers.push_back(matches[2 * i + 1]); }
  if (o.detect_watermark && num > 0) {  // DetectWatermark
    const double d1 = o.watermark_border_size * std::hypot(c1.w, c1.h), d2 = o.watermark_border_size * std::hypot(c2.w, c2.h);
    std::vector<double> a, b, c, d;
    int border = 0;
    for (int64_t i = 0; i < m; ++i) {
      if (!(*mask)[i]) continue;
      a.push_back(x1[i]); b.push_back(y1[i]); c.push_back(x2[i]); d.push_back(y2[i]);
      const bool b1 = x1[i] < d1 || x1[i] > c1.w - d1 || y1[i] < d1 || y1[i] > c1.h - d1;
      const bool b2 = x2[i] < d2 || x2[i] > c2.w - d2 || y2[i] < d2 || y2[i] > c2.h - d2;
      border += (b1 && b2);
    }
    if (static_cast<double>(border) / num >= o.watermark_min_inlier_ratio) {
      const Report t = loransac(Problem{3, &a, &b, &c, &d}, o, o.max_error, o.watermark_min_inlier_ratio, rng);
      if (t.success && static_cast<double>(t.num_inliers) / num >= o.watermark_min_inlier_ratio) R.config = WATERMARK;
    }
  }
  return R;
}

Opts unpack(const double* v) {
  Opts o;
  o.min_num_inliers = static_cast<int>(v[0]); o.min_E_F_inlier_ratio = v[1]; o.max_H_inlier_ratio = v[2];
  o.watermark_min_inlier_ratio = v[3]; o.watermark_border_size = v[4]; o.detect_watermark = static_cast<int>(v[5]);
  o.force_H_use = static_cast<int>(v[6]); o.max_error = v[7]; o.min_inlier_ratio = v[8]; o.confidence = v[9];
  o.dyn_mult = v[10]; o.min_num_trials = static_cast<int>(v[11]); o.max_num_trials = static_cast<int>(v[12]);
  return o;
}
Cam unpack_cam(const double* v) {
  return Cam{v[0], v[1], v[2], v[3], static_cast<int>(v[4]), static_cast<int>(v[5]), static_cast<int>(v[6])};
}

}  // namespace

extern "C" {

// cam: fx fy cx cy width height has_prior (pinhole models only); opts: 13 doubles (see unpack).
// out_i32: config nE nF nH n_inliers; out_inliers: capacity m x 2; out_models: E F H (27 doubles); models_scored.
int orc_estimate_two_view_geometry(const double* cam1, const double* pts1, const double* cam2, const double* pts2,
                                   const uint32_t* matches, int64_t m, const double* opts, uint32_t seed, int32_t* out_i32,
                                   uint32_t* out_inliers, double* out_models, int64_t* models_scored) {
  const Result R = estimate_tvg(unpack_cam(cam1), pts1, unpack_cam(cam2), pts2, matches, m, unpack(opts), seed);
  out_i32[0] = R.config; out_i32[1] = R.nE; out_i32[2] = R.nF; out_i32[3] = R.nH;
  out_i32[4] = static_cast<int32_t>(R.inliers.size() / 2);
  if (out_inliers && !R.inliers.empty()) std::memcpy(out_inliers, R.inliers.data(), R.inliers.size() * sizeof(uint32_t));
  if (out_models) {
    std::memcpy(out_models, R.E, 72); std::memcpy(out_models + 9, R.F, 72); std::memcpy(out_models + 18, R.H, 72);
  }
  if (models_scored) *models_scored = R.models_scored;
  return 0;
}

// n_jobs pairs on n_threads host threads (a pair per thread at a time).  Job k: cameras cam1[k] / cam2[k] (7 doubles
// each), keypoint arrays pts1[k] / pts2[k] ([.. x 2] doubles), matches[k] ([m[k] x 2]).  out_i32: [n_jobs][5].
int orc_verify_pairs(int64_t n_jobs, const double* cams1, const double* const* pts1, const double* cams2,
                     const double* const* pts2, const uint32_t* const* matches, const int64_t* m, const double* opts,
                     uint32_t seed, int n_threads, int32_t* out_i32, int64_t* models_scored) {
  std::vector<std::thread> pool;
  std::vector<long> scored(std::max(1, n_threads), 0);
  auto work = [&](int t) {
    for (int64_t k = t; k < n_jobs; k += n_threads) {
      int64_t ms = 0;
      orc_estimate_two_view_geometry(cams1 + 7 * k, pts1[k], cams2 + 7 * k, pts2[k], matches[k], m[k], opts,
                                     seed + static_cast<uint32_t>(k), out_i32 + 5 * k, nullptr, nullptr, &ms);
      scored[t] += ms;
    }
  };
  n_threads = std::max(1, n_threads);
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (std::thread& th : pool) th.join();
  if (models_scored) {
    *models_scored = 0;
    for (long v : scored) *models_scored += v;
  }
  return 0;
}

}  // extern "C"
