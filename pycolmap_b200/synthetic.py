"""Seeded synthetic inputs for tests and bench.py.

SIFT-like descriptors follow the upstream test recipe (U:feature/sift_test.cc
CreateRandomFeatureDescriptors): g = U(0,1)^2 per element, L2-normalise, round(512 g) saturated
to uint8 -- so that dot products of matching descriptors sit near 2^18 and the acos/ratio tests
behave as on real SIFT.  Uniform random bytes would clamp every distance to acos(1) = 0.
"""
import numpy as np


def sift_like(rng, n):
    g = rng.random((n, 128)) ** 2
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    return np.clip(np.round(512.0 * g), 0, 255).astype(np.uint8)


def perturb(rng, d, sigma=4.0):
    """Noisy copy of descriptors (a re-observation of the same 3-D point)."""
    x = d.astype(np.float64) + rng.normal(0.0, sigma, d.shape)
    x = np.clip(x, 0, None)
    x *= 512.0 / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-9)
    return np.clip(np.round(x), 0, 255).astype(np.uint8)


def matching_pair(rng, n1, n2, n_common, sigma=4.0):
    """Two descriptor sets sharing `n_common` (perturbed) descriptors at random positions."""
    d1 = sift_like(rng, n1)
    d2 = sift_like(rng, n2)
    n_common = min(n_common, n1, n2)
    i1 = rng.choice(n1, n_common, replace=False)
    i2 = rng.choice(n2, n_common, replace=False)
    d2[i2] = perturb(rng, d1[i1], sigma)
    return d1, d2, np.stack([i1, i2], 1)


def exhaustive_pairs(n):
    """All unordered pairs (i < j), the pair SET of ExhaustiveFeatureMatcher (row P1)."""
    i, j = np.triu_indices(n, 1)
    return np.stack([i, j], 1).astype(np.int32)


# ---------------------------------------------------------------------------------------------
# Synthetic multi-view scene (torch; runs on CPU or CUDA).  SURVEY.md section 8(d).
# ---------------------------------------------------------------------------------------------

def make_scene(n_images, n_feat, seed=0, device="cpu", window_images=24.0, vis_frac=0.45, detect_prob=0.7,
               desc_sigma=5.0, px_sigma=0.5, width=1600, height=1200, focal=1200.0, image_range=None):
    """Cameras on a circle looking at a rough cylindrical surface; SIMPLE_PINHOLE.

    Image i observes the surface points whose azimuth lies within +-w of the camera azimuth
    (w = window_images * 2*pi / n_images), each detected with probability `detect_prob`; the
    observations (noisy projections + re-normalised noisy copies of the point's base descriptor)
    are padded to exactly `n_feat` rows with distractors (fresh SIFT-like descriptors, uniform
    keypoints).  Image pairs further apart than ~2*window_images share nothing.

    image_range=(lo, hi) generates only that slice of images (same world for every slice; used by
    the multi-GPU bench where each rank produces its own shard before the NCCL all-gather).

    Returns dict(desc uint8 [n,K,128], kpts float32 [n,K,2], point_id int64 [n,K] (-1 = distractor),
                 cameras list of dict).
    """
    import math
    import torch

    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    N, K = n_images, n_feat
    w = window_images * 2.0 * math.pi / N
    n_vis_target = int(vis_frac * K / detect_prob)
    P = max(16, int(n_vis_target * math.pi / w))  # a 2w window holds ~n_vis_target points
    r0, R = 4.0, 10.0
    phi = torch.rand(P, generator=g, device=dev) * (2 * math.pi)
    phi, _ = torch.sort(phi)
    rad = r0 + 0.8 * (torch.rand(P, generator=g, device=dev) - 0.5)
    zz = 3.0 * (torch.rand(P, generator=g, device=dev) - 0.5)
    X = torch.stack([rad * torch.cos(phi), rad * torch.sin(phi), zz], 1)  # [P,3]
    base = torch.rand(P, 128, generator=g, device=dev) ** 2
    base = base / base.norm(dim=1, keepdim=True) * 512.0

    lo_i, hi_i = image_range if image_range is not None else (0, N)
    n_out = hi_i - lo_i
    desc = torch.empty(n_out, K, 128, dtype=torch.uint8, device=dev)
    kpts = torch.empty(n_out, K, 2, dtype=torch.float32, device=dev)
    pid = torch.full((n_out, K), -1, dtype=torch.int64, device=dev)
    cx, cy = width / 2.0, height / 2.0
    phi_c = phi.cpu()

    def bound(x):
        return int(torch.searchsorted(phi_c, torch.tensor([x])).item())

    for i in range(lo_i, hi_i):
        gi = torch.Generator(device=dev)
        gi.manual_seed(seed * 1000003 + i + 1)  # per-image stream: independent of sharding
        o = i - lo_i
        th = 2 * math.pi * i / N
        C = torch.tensor([R * math.cos(th), R * math.sin(th), 0.0], device=dev)
        zc = -C / C.norm()                     # optical axis towards the origin
        xc = torch.tensor([-math.sin(th), math.cos(th), 0.0], device=dev)
        yc = torch.linalg.cross(zc, xc)
        Rm = torch.stack([xc, yc, zc], 0)      # world -> camera
        idx = torch.arange(bound(th - w), bound(th + w), device=dev)
        if th - w < 0:
            idx = torch.cat([torch.arange(bound(th - w + 2 * math.pi), P, device=dev), idx])
        if th + w > 2 * math.pi:
            idx = torch.cat([idx, torch.arange(0, bound(th + w - 2 * math.pi), device=dev)])
        keep = torch.rand(len(idx), generator=gi, device=dev) < detect_prob
        idx = idx[keep]
        Xc = (X[idx] - C) @ Rm.T
        u = focal * Xc[:, 0] / Xc[:, 2] + cx
        v = focal * Xc[:, 1] / Xc[:, 2] + cy
        inside = (Xc[:, 2] > 0.1) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
        idx, u, v = idx[inside], u[inside], v[inside]
        nv = min(len(idx), K)
        if len(idx) > nv:
            sel = torch.randperm(len(idx), generator=gi, device=dev)[:nv]
            idx, u, v = idx[sel], u[sel], v[sel]
        d = base[idx] + desc_sigma * torch.randn(nv, 128, generator=gi, device=dev)
        d = torch.clamp(d, min=0)
        d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-9) * 512.0
        d = torch.clamp(torch.round(d), 0, 255).to(torch.uint8)
        perm = torch.randperm(K, generator=gi, device=dev)
        rows_v, rows_d = perm[:nv], perm[nv:]
        desc[o, rows_v] = d
        kpts[o, rows_v, 0] = u + px_sigma * torch.randn(nv, generator=gi, device=dev)
        kpts[o, rows_v, 1] = v + px_sigma * torch.randn(nv, generator=gi, device=dev)
        pid[o, rows_v] = idx
        nd = K - nv
        if nd > 0:
            x = torch.rand(nd, 128, generator=gi, device=dev) ** 2
            x = x / x.norm(dim=1, keepdim=True) * 512.0
            desc[o, rows_d] = torch.clamp(torch.round(x), 0, 255).to(torch.uint8)
            kpts[o, rows_d, 0] = torch.rand(nd, generator=gi, device=dev) * width
            kpts[o, rows_d, 1] = torch.rand(nd, generator=gi, device=dev) * height
    cams = [dict(model=0, width=width, height=height, params=[focal, cx, cy], has_prior_focal_length=1)
            for _ in range(n_out)]
    return dict(desc=desc, kpts=kpts, point_id=pid, cameras=cams)
