#!/usr/bin/env python
"""DB-inclusive end to end: the plugin call a pycolmap user makes -- pycolmap_b200.match_exhaustive(database_path) on a
COLMAP-schema SQLite database -- timed by the wall clock, with the pipeline's own breakdown (read, upload, GPU, write on
the writer thread, wait for the writer).  SURVEY.md section 7 item 7 / BASELINE configs[0] (50 x 4096) and a larger one.

    python tools/bench_db.py --images 50 --feats 4096
    python tools/bench_db.py --images 1000 --feats 8192 --out gpurun_out/db_1000.json
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=50)
    ap.add_argument("--feats", type=int, default=4096)
    ap.add_argument("--block-size", type=int, default=50)
    ap.add_argument("--gpu-index", default="0")
    ap.add_argument("--dir", default=None, help="where the database goes (default: a temporary directory, /dev/shm if present)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import pycolmap_b200 as pb
    from pycolmap_b200 import synthetic as syn

    base = args.dir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    tmp = tempfile.mkdtemp(prefix="b2m_db_", dir=base)
    path = os.path.join(tmp, "scene.db")
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    t0 = time.perf_counter()
    scene = syn.make_scene(args.images, args.feats, seed=0, device=dev)
    desc, kpts = scene["desc"].cpu().numpy(), scene["kpts"].cpu().numpy()
    with pb.Database(path) as db:
        cid = db.add_camera(0, 1600, 1200, [1200.0, 800.0, 600.0], True)
        db.begin()
        for i in range(args.images):
            iid = db.add_image(f"frame{i:05d}.png", cid)
            kp = np.zeros((args.feats, 6), np.float32)
            kp[:, :2] = kpts[i]
            db.write_keypoints(iid, kp)
            db.write_descriptors(iid, desc[i])
        db.commit()
    build_s = time.perf_counter() - t0
    size_in = os.path.getsize(path)
    pb.Context(device=0).close()                      # CUDA context / module load outside the timed call
    t0 = time.perf_counter()
    pb.match_exhaustive(path, sift_options={"gpu_index": args.gpu_index}, matching_options={"block_size": args.block_size})
    wall = time.perf_counter() - t0
    tm = pb.last_pipeline_timing()
    n_pairs = args.images * (args.images - 1) // 2
    with pb.Database(path) as db:
        verified, n_m = db.num_verified_image_pairs, db.num_matches
    t0 = time.perf_counter()
    pb.match_exhaustive(path, sift_options={"gpu_index": args.gpu_index}, matching_options={"block_size": args.block_size})
    resume = time.perf_counter() - t0
    out = {"what": "pycolmap_b200.match_exhaustive(database_path): database read + upload + match + verify + database write",
           "images": args.images, "features_per_image": args.feats, "pairs": n_pairs, "verified_pairs": verified,
           "matches_written": n_m, "wall_s": wall, "pairs_per_s_db_inclusive": n_pairs / wall, "breakdown": tm,
           "resume_noop_s": resume, "db_bytes_before": size_in, "db_bytes_after": os.path.getsize(path),
           "db_location": tmp, "build_db_s": build_s, "gpu_index": args.gpu_index}
    print(json.dumps(out))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    os.remove(path)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
