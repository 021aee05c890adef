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
