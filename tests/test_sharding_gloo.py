"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py -- the launcher's side channel carries a
communicator id from rank 0 to every rank, every rank builds ONLY the images b2m_comm_image_range gives it, one
all-gather (gloo here; NCCL inside b2m_set_images_sharded on GPUs) reproduces the full descriptor set, the pair
list dealt round-robin partitions the exhaustive list, and per-rank oracle matching concatenates to the
single-process result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import pycolmap_b200 as pb
    from pycolmap_b200 import synthetic as syn
    box = [os.urandom(128) if rank == 0 else None]        # stands for Context.comm_unique_id() (needs NCCL + a GPU)
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], bytes) and len(box[0]) == 128
    n_img, K = 7, 256                                      # 7 images over 2 ranks: ragged shards (4 + 3)
    first, count = pb.comm_image_range(n_img, world, rank)
    per = pb.comm_image_range(n_img, world, 0)[1]
    scene = syn.make_scene(n_img, K, seed=3, window_images=2.0, image_range=(first, first + count))
    local = scene["desc"].reshape(-1, 128)
    pad = (per - count) * K
    if pad:
        local = torch.cat([local, torch.zeros((pad, 128), dtype=local.dtype)])
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())
    full = torch.cat(parts)[: n_img * K]
    ref = syn.make_scene(n_img, K, seed=3, window_images=2.0)["desc"].reshape(-1, 128)
    assert torch.equal(full, ref), "all-gathered set differs from the single-process set"
    pairs = np.concatenate(pb.exhaustive_pair_blocks(n_img, 3))
    mine = np.ascontiguousarray(pairs[rank::world])
    res = oracle.fast_match_pairs(full.numpy(), np.full(n_img, K, np.int32), mine, n_threads=2)
    np.save(os.path.join(out_dir, f"pairs{rank}.npy"), mine)
    np.save(os.path.join(out_dir, f"counts{rank}.npy"), np.array([len(m) for m in res]))
    np.save(os.path.join(out_dir, f"id{rank}.npy"), np.frombuffer(box[0], np.uint8))
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t)
    assert int(t.item()) == len(pairs)
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    import oracle
    import pycolmap_b200 as pb
    from pycolmap_b200 import synthetic as syn
    pairs = np.concatenate(pb.exhaustive_pair_blocks(7, 3))
    p0, p1 = np.load(tmp_path / "pairs0.npy"), np.load(tmp_path / "pairs1.npy")
    assert np.array_equal(np.load(tmp_path / "id0.npy"), np.load(tmp_path / "id1.npy"))   # the id reached rank 1
    got = {tuple(sorted(p)) for p in np.concatenate([p0, p1]).tolist()}
    assert got == {tuple(sorted(p)) for p in pairs.tolist()} and len(p0) + len(p1) == len(pairs) == 21
    desc = syn.make_scene(7, 256, seed=3, window_images=2.0)["desc"].reshape(-1, 128).numpy()
    want = oracle.fast_match_pairs(desc, np.full(7, 256, np.int32), pairs, n_threads=2)
    want_counts = {tuple(p): len(m) for p, m in zip(pairs.tolist(), want)}
    for r, pr in enumerate((p0, p1)):
        counts = np.load(tmp_path / f"counts{r}.npy")
        for p, c in zip(pr.tolist(), counts):
            assert want_counts[tuple(p)] == c


def test_image_ranges_and_pair_dealing():
    sys.path.insert(0, ROOT)
    import pycolmap_b200 as pb
    for n, w in [(1000, 8), (7, 2), (5, 8), (1, 1), (1414, 2), (0, 4)]:
        cover, per = [], -(-n // w)
        for r in range(w):
            lo, cnt = pb.comm_image_range(n, w, r)
            cover += list(range(lo, lo + cnt))
            assert 0 <= cnt <= per and (cnt == per or lo + cnt == n)
        assert cover == list(range(n))
        pairs = np.arange(2 * 37).reshape(-1, 2)
        parts = [pairs[r::w] for r in range(w)]
        assert sum(len(p) for p in parts) == len(pairs)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
