#!/usr/bin/env python
"""GPUTEST_r01 root cause: Python mirror vs C++ host on the same distorted-camera estimator call."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ransac as R
import pycolmap_b200 as pb
import pycolmap_b200.native as nat
from helpers import scenes
cam = dict(model=7, params=[1200.0, 1200.0, 800.0, 600.0, 0.7], width=1600, height=1200, has_prior_focal_length=1)
rng = np.random.default_rng(5)
p1, p2, planted = scenes.two_view_scene(rng, 400, 0.3, "general")
d1 = R.img_from_cam(cam, (p1 - [800.0, 600.0]) / 1200.0)
d2 = R.img_from_cam(cam, (p2 - [800.0, 600.0]) / 1200.0)
print("d1 flags", d1.flags["C_CONTIGUOUS"], d1.dtype, d1.shape, d1.strides)
g = nat.estimate_two_view_geometry(cam, d1, cam, d2)
gp = pb.estimate_two_view_geometry(cam, d1, cam, d2)
gc = nat.estimate_two_view_geometry(cam, np.ascontiguousarray(d1), cam, np.ascontiguousarray(d2))
for name, x in (("native", g), ("mirror", gp), ("native_contig", gc)):
    im = np.asarray(x.inlier_matches)
    print(name, int(x.config), len(im), im[:5].tolist(), np.asarray(x.E).ravel()[:3])
a, b = np.asarray(g.inlier_matches), np.asarray(gp.inlier_matches)
print("shapes", a.shape, b.shape, a.dtype, b.dtype, "equal", np.array_equal(a, b))
if a.shape == b.shape:
    print("first diffs", np.argwhere(a != b)[:5].tolist())
print("mirror nEFH", getattr(gp, "num_inliers_EFH", None), "native nEFH", g.num_inliers_EFH)
