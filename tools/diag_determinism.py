#!/usr/bin/env python
"""Run-to-run / context-to-context determinism of the stand-alone verifier with a distorted camera
(the GPUTEST_r01 failure: two hosts, same ABI call, same seed, different inlier lists).
Prints one line per repetition; exit code 1 when two repetitions differ."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ransac as R  # noqa: E402
import pycolmap_b200 as nat  # noqa: E402
from helpers import scenes  # noqa: E402

n = int(os.environ.get("DIAG_N", "400"))
reps = int(os.environ.get("DIAG_REPS", "4"))
cams = {
    "PINHOLE0": dict(scenes.CAM),
    "FOV": dict(model=7, params=[1200.0, 1200.0, 800.0, 600.0, 0.7], width=1600, height=1200, has_prior_focal_length=1),
    "OPENCV": dict(model=4, params=[1200.0, 1200.0, 800.0, 600.0, -0.12, 0.03, 1e-3, -2e-3], width=1600, height=1200,
                   has_prior_focal_length=1),
}
rng = np.random.default_rng(5)
p1, p2, planted = scenes.two_view_scene(rng, n, 0.3, "general")
bad = 0
for name, cam in cams.items():
    if name == "PINHOLE0":
        d1, d2 = p1, p2
    else:
        d1 = R.img_from_cam(cam, (p1 - [800.0, 600.0]) / 1200.0)
        d2 = R.img_from_cam(cam, (p2 - [800.0, 600.0]) / 1200.0)
    ref = None
    for ctx_i in range(2):
        ctx = nat.Context(device=0, seed=0)
        for r in range(reps):
            g = ctx.estimate_two_view_geometry(cam, d1, cam, d2)
            key = (int(g.config), tuple(g.num_inliers_EFH), g.inlier_matches.tobytes(), np.asarray(g.E).tobytes())
            if ref is None:
                ref = key
            same = key == ref
            bad += (not same)
            print(f"{name} ctx{ctx_i} rep{r}: config={int(g.config)} nEFH={tuple(g.num_inliers_EFH)} "
                  f"n_inl={len(g.inlier_matches)} same_as_first={same}", flush=True)
        ctx.close()
print("DIAG", "MISMATCH" if bad else "deterministic", bad)
sys.exit(1 if bad else 0)
