"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py / sharding.py -- every rank builds its
image shard, one all-gather reproduces the full descriptor set, the dealt pair lists partition the
exhaustive list, and per-rank oracle matching of the shards concatenates to the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from pycolmap_b200 import sharding, synthetic as syn
    n_img, K = 7, 256                      # 7 images over 2 ranks: ragged shards (4 + 3)
    lo, hi, per = sharding.image_shard(n_img, rank, world)
    scene = syn.make_scene(n_img, K, seed=3, window_images=2.0, image_range=(lo, hi))
    full = sharding.all_gather_rows(scene["desc"].reshape(-1, 128), n_img, K, rank, world, dist)
    ref = syn.make_scene(n_img, K, seed=3, window_images=2.0)["desc"].reshape(-1, 128)
    assert torch.equal(full, ref), "all-gathered set differs from the single-process set"
    pairs = syn.exhaustive_pairs(n_img)
    mine = sharding.pair_shard(pairs, rank, world)
    res = oracle.fast_match_pairs(full.numpy(), np.full(n_img, K, np.int32), mine, n_threads=2)
    np.save(os.path.join(out_dir, f"pairs{rank}.npy"), mine)
    np.save(os.path.join(out_dir, f"counts{rank}.npy"), np.array([len(m) for m in res]))
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t)
    assert int(t.item()) == len(pairs)
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    import oracle
    from pycolmap_b200 import synthetic as syn
    pairs = syn.exhaustive_pairs(7)
    p0, p1 = np.load(tmp_path / "pairs0.npy"), np.load(tmp_path / "pairs1.npy")
    got = {tuple(p) for p in np.concatenate([p0, p1]).tolist()}
    assert got == {tuple(p) for p in pairs.tolist()} and len(p0) + len(p1) == len(pairs)
    desc = syn.make_scene(7, 256, seed=3, window_images=2.0)["desc"].reshape(-1, 128).numpy()
    want = oracle.fast_match_pairs(desc, np.full(7, 256, np.int32), pairs, n_threads=2)
    want_counts = {tuple(p): len(m) for p, m in zip(pairs.tolist(), want)}
    for r, pr in enumerate((p0, p1)):
        counts = np.load(tmp_path / f"counts{r}.npy")
        for p, c in zip(pr.tolist(), counts):
            assert want_counts[tuple(p)] == c


def test_shard_arithmetic():
    from pycolmap_b200 import sharding
    for n, w in [(1000, 8), (7, 2), (5, 8), (1, 1), (1414, 2)]:
        cover = []
        for r in range(w):
            lo, hi, per = sharding.image_shard(n, r, w)
            cover += list(range(lo, hi))
            assert hi - lo <= per
        assert cover == list(range(n))
        pairs = np.arange(2 * 37).reshape(-1, 2)
        parts = [sharding.pair_shard(pairs, r, w) for r in range(w)]
        assert sum(len(p) for p in parts) == len(pairs)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
