#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference (/root/reference) has NO golden vectors for this path and cannot be run here (it needs
COLMAP 3.9.1, un-vendored) -- parity is unpinned.  These fixtures therefore pin
the ORACLE's literal restatement (oracle_match.c orc_match_bruteforce / orc_match_guided: materialised
distance matrix, row scan, column scan) on small seeded inputs, so that neither the oracle's streaming
SIMD path nor the CUDA path can drift unnoticed.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pycolmap_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260922)
    cases = {}
    for name, (n1, n2, common) in {"a": (37, 53, 20), "b": (200, 130, 90), "c": (260, 515, 200)}.items():
        d1, d2, _ = syn.matching_pair(rng, n1, n2, common)
        if name == "b":                       # duplicates, zero rows, exact ties
            d2[5] = d2[7]
            d1[3] = 0
            d2[11] = 0
            d1[20:23] = d1[20]
        cases[f"{name}_d1"], cases[f"{name}_d2"] = d1, d2
        for tag, kw in {"default": {}, "nocross": {"cross_check": False},
                        "loose": {"max_ratio": 1.0, "max_distance": float(np.pi)},
                        "tight": {"max_ratio": 0.6, "max_distance": 0.5}}.items():
            cases[f"{name}_{tag}"] = oracle.match_bruteforce(d1, d2, **kw)
    # guided matching under a homography
    n = 150
    d1 = syn.sift_like(rng, n)
    d2 = syn.perturb(rng, d1)
    kp1 = rng.uniform(0, 1000, (n, 2)).astype(np.float32)
    H = np.array([[1.0, 0.02, 5.0], [-0.01, 1.0, -3.0], [1e-5, 0, 1.0]])
    p = np.c_[kp1, np.ones(n)] @ H.T
    kp2 = (p[:, :2] / p[:, 2:]).astype(np.float32)
    kp2[:40] += 50.0
    cases.update(g_d1=d1, g_d2=d2, g_kp1=kp1, g_kp2=kp2, g_H=H,
                 g_matches=oracle.match_guided(d1, kp1, d2, kp2, 1, H, 4.0))
    np.savez_compressed(os.path.join(HERE, "match_golden.npz"), **cases)
    print("wrote", os.path.join(HERE, "match_golden.npz"), {k: v.shape for k, v in cases.items() if "_d" not in k})


if __name__ == "__main__":
    main()
