"""Golden fixtures (tests/golden/match_golden.npz, made by tests/golden/make_golden.py from the oracle's
literal restatement; the reference itself holds none): the streaming-SIMD oracle on CPU, and the CUDA
path on GPU, must reproduce them bit for bit."""
import os

import numpy as np
import pytest

import oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "match_golden.npz"))
CASES = [(n, t, kw) for n in "abc" for t, kw in (("default", {}), ("nocross", {"cross_check": False}),
                                                   ("loose", {"max_ratio": 1.0, "max_distance": float(np.pi)}),
                                                   ("tight", {"max_ratio": 0.6, "max_distance": 0.5}))]


@pytest.mark.parametrize("name,tag,kw", CASES)
def test_oracle_fast_path_reproduces_golden(name, tag, kw):
    assert np.array_equal(oracle.fast_match_pair(G[f"{name}_d1"], G[f"{name}_d2"], **kw), G[f"{name}_{tag}"])
    assert np.array_equal(oracle.match_bruteforce(G[f"{name}_d1"], G[f"{name}_d2"], **kw), G[f"{name}_{tag}"])


def test_oracle_guided_reproduces_golden():
    m = oracle.match_guided(G["g_d1"], G["g_kp1"], G["g_d2"], G["g_kp2"], 1, G["g_H"], 4.0)
    assert np.array_equal(m, G["g_matches"]) and set(m[:, 0]) == set(range(40, 150))


@pytest.mark.gpu
@pytest.mark.parametrize("name,tag,kw", CASES)
def test_gpu_reproduces_golden(ctx, name, tag, kw):
    assert np.array_equal(ctx.match_pair(G[f"{name}_d1"], G[f"{name}_d2"], dict(kw)), G[f"{name}_{tag}"])
