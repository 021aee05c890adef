#!/usr/bin/env python
"""bench.py -- image-pairs matched(+verified)/s on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 2 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference          # the reference's CPU algorithm on the host cores

A "step" is one pass of the hot path (exhaustive matching [+ two-view verification]) over every
image pair of the synthetic scene.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def log(msg):
    """Progress on stderr (stdout carries exactly one JSON line)."""
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=1000, help="images at 1 GPU (scaled by sqrt(gpus): weak scaling)")
    ap.add_argument("--feats", type=int, default=8192)
    ap.add_argument("--verify", type=int, default=-1, help="1: match + two-view verification, 0: match only")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pair-batch", type=int, default=0)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _verify_worker(job):
    """One pair through the sequential LO-RANSAC oracle (a verifier thread of the reference)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import ransac as R
    cam, kp1, kp2, m, seed = job
    t0 = time.perf_counter()
    g = R.estimate_two_view_geometry(cam, kp1, cam, kp2, m, seed=seed)
    return time.perf_counter() - t0, int(g.config), len(g.inlier_matches)


def cpu_baseline(desc_np, n_feat, pairs, budget_s, verify, kpts_np=None, cam=None, full_frac=None):
    """Oracle (CPU port of the reference algorithm) on a bounded sample of the same workload,
    all host threads (one pair per thread, like upstream's FeatureMatcherWorker / VerifierWorker
    pools).  Matching: the AVX-512-VNNI brute-force matcher.  Verification (when `verify`): the
    sequential numpy LO-RANSAC oracle on up to 2 x cores of the sampled pairs that have >= 15 matches;
    pairs/s = cores / (core-seconds per matched pair + verified fraction x core-seconds per verification)."""
    import oracle
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(123)
    probe = pairs[rng.choice(len(pairs), min(len(pairs), 2 * cores), replace=False)]
    t0 = time.perf_counter()
    oracle.fast_match_pairs(desc_np, n_feat, probe, n_threads=cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    rate = len(probe) / dt
    n = int(min(len(pairs), max(2 * cores, rate * budget_s * 0.7)))
    sample = pairs[rng.choice(len(pairs), n, replace=False)]
    t0 = time.perf_counter()
    res = oracle.fast_match_pairs(desc_np, n_feat, sample, n_threads=cores)
    dt = time.perf_counter() - t0
    out = {"value": n / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
           "match_pairs_per_s": n / dt,
           "sample": f"{n} random pairs of the same scene, oracle.fast_match_pairs ({oracle.fast_isa()}), {dt:.1f} s"}
    if verify and kpts_np is not None:
        import multiprocessing as mp
        K = int(n_feat[0])
        cand = [k for k in range(n) if len(res[k]) >= 15]
        frac = len(cand) / max(n, 1)
        cand = cand[: 2 * cores]
        if cand:
            jobs = [(cam, kpts_np[sample[k, 0] * K:(sample[k, 0] + 1) * K].astype(np.float64),
                     kpts_np[sample[k, 1] * K:(sample[k, 1] + 1) * K].astype(np.float64), res[k], k) for k in cand]
            procs = min(cores, len(jobs), 64)
            jobs = jobs[: 2 * procs]
            t0 = time.perf_counter()
            # "spawn": the parent may hold a CUDA context and helper threads -- fork() is not safe there.
            # One BLAS thread per worker: the children must see these BEFORE they import numpy (128
            # processes x 128 OpenBLAS threads on 9x9 matrices made this leg crawl for minutes).
            saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
            os.environ.update({k: "1" for k in saved})
            try:
                with mp.get_context("spawn").Pool(procs) as pool:
                    pool.map_async(_verify_worker, jobs[:procs]).get(timeout=120)   # warm-up: interpreter + imports
                    t0 = time.perf_counter()
                    pool.map_async(_verify_worker, jobs).get(timeout=180)
            except Exception as e:  # noqa: BLE001  (never let the baseline leg hang the bench)
                out["verify_error"] = repr(e)
                return out, sample, res
            finally:
                for k, v in saved.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
            wall = time.perf_counter() - t0
            core_s_verify = wall * procs / len(jobs)
            core_s_match = cores * dt / n
            out["verify_core_seconds_per_pair"] = core_s_verify
            # the sample comes from the first images of the scene (denser in overlapping pairs than the
            # whole exhaustive set): weight with the verified fraction of the FULL workload when known
            use = frac if full_frac is None else full_frac
            out["verified_fraction_of_pairs"] = use
            out["verified_fraction_in_sample"] = frac
            out["value"] = cores / (core_s_match + use * core_s_verify)
            out["sample"] += (f"; + oracle.ransac (numpy, sequential LO-RANSAC) on {len(jobs)} of the sampled pairs with "
                              f">= 15 matches ({procs} processes, {wall:.1f} s)")
    return out, sample, res


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    verify = args.verify if args.verify >= 0 else 1
    n_img = int(round(args.images * math.sqrt(max(world, 1))))
    K = args.feats
    workload = (f"{n_img} images x {K} SIFT-like uint8 128-D descriptors, exhaustive matching"
                + (" + two-view verification (E/F/H LO-RANSAC)" if verify else "")
                + f"; BASELINE configs[{2 if verify else 1}] scaled to {n_img} images for {world} GPU(s)")
    metric = "image-pairs matched+verified/sec" if verify else "image-pairs matched/sec"

    from pycolmap_b200 import synthetic as syn

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's own CPU path cannot be built here (pycolmap -> COLMAP 3.9.1, un-vendored):
        # this arm times the oracle port of its algorithm on the host cores.
        n_small = min(n_img, 64)
        scene = syn.make_scene(n_img, K, seed=0, device="cpu", image_range=(0, n_small))
        desc = scene["desc"].numpy().reshape(-1, 128)
        nf = np.full(n_small, K, np.int32)
        pairs = syn.exhaustive_pairs(n_small)
        vals, step_ms = [], []
        for it in range(args.warmup + args.steps):
            t_step = time.perf_counter()
            # verified fraction of the full exhaustive workload from the scene geometry: images further
            # apart than 2 x window_images (default 24) share no points (pycolmap_b200/synthetic.py)
            full_frac = min(1.0, 2.0 * (2 * 24 - 1) / max(n_img - 1, 1))
            cb, _, _ = cpu_baseline(desc, nf, pairs, max(2.0, args.cpu_seconds / 2), verify,
                                    scene["kpts"].numpy().reshape(-1, 2), scene["cameras"][0], full_frac=full_frac)
            if it >= args.warmup:
                vals.append(cb)
                step_ms.append((time.perf_counter() - t_step) * 1e3)
        v = float(np.mean([c["value"] for c in vals])) if vals else 0.0
        cb = vals[-1] if vals else {"cores": os.cpu_count(), "kind": "port", "sample": "none"}
        cb["value"] = v
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(step_ms)) if step_ms else None,  # wall time of one bounded sample
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "note": "first 64 images of the same scene; bounded sample per step"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    import pycolmap_b200 as pb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: every rank synthesises its own shard of images, ONE all-gather makes the set resident
    from pycolmap_b200 import sharding
    lo, hi, _per = sharding.image_shard(n_img, rank, world)
    scene = syn.make_scene(n_img, K, seed=0, device=dev, image_range=(lo, hi))
    desc_full = sharding.all_gather_rows(scene["desc"].reshape(-1, 128), n_img, K, rank, world, dist)
    kpts_full = sharding.all_gather_rows(scene["kpts"].reshape(-1, 2), n_img, K, rank, world, dist)
    desc_full, kpts_full = desc_full.contiguous(), kpts_full.contiguous()
    torch.cuda.synchronize()
    if rank == 0:
        log(f"scene ready: {n_img} images x {K} features on {world} rank(s)")
    cams = [dict(model=0, width=1600, height=1200, params=[1200.0, 800.0, 600.0], has_prior_focal_length=1)
            for _ in range(n_img)]
    nfeat = np.full(n_img, K, np.int32)

    # every unordered pair once, visited block by block like ExhaustiveFeatureMatcher::Run with the
    # default block_size = 50 (U:controllers/feature_matching.cc): a 50 x 50 block re-uses 100 images
    # (105 MB of descriptors), which stay L2-resident
    from pycolmap_b200.pipeline import exhaustive_pair_blocks
    all_pairs = np.concatenate(list(exhaustive_pair_blocks(n_img, 50)))
    my_pairs = sharding.pair_shard(all_pairs, rank, world)       # independent units, no data-path collective

    ctx = pb.Context(device=local_rank, pair_batch=args.pair_batch)
    ctx.set_images_device(nfeat, desc_full.data_ptr(), kpts_full.data_ptr() if verify else None,
                          cams if verify else None)
    sift = ctx.sift_opts()
    tvg = ctx.tvg_opts() if verify else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        res = ctx.match_pairs(my_pairs, sift, tvg)
        st = ctx.stats()
        out = (st.last_total_ms, st.last_k1_ms, st.last_k1_launches, res.total_matches, st.last_verify_ms,
               res.num_verified)
        res.free()
        return out

    for _ in range(args.warmup):
        one_step()
    barrier()
    if rank == 0:
        log("warm-up done")
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ctx.stats().kernel_launches
    t_wall0 = time.perf_counter()
    dev_ms, k1_ms, k1_n, total_matches, ver_ms = 0.0, 0.0, 0, 0, 0.0
    for _ in range(args.steps):
        a, b, c, d, e, n_ver = one_step()
        ver_ms += e
        dev_ms += a
        k1_ms += b
        k1_n += c
        total_matches = d
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    if rank == 0:
        log(f"timed region done: {wall_ms / max(args.steps, 1):.0f} ms/step")
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.stats().kernel_launches - launches0

    t = torch.tensor([dev_ms, wall_ms, k1_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(len(my_pairs)), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, k1_ms_max = t.tolist()
    pairs_total, launches_total = cnt.tolist()
    verified_fraction = (n_ver / max(len(my_pairs), 1)) if verify else 0.0
    ms_per_step = dev_ms_max / args.steps
    value = pairs_total / (ms_per_step / 1e3)

    # ---- e2e: same call with HOST buffers; H2D of the descriptor set and D2H of the results inside
    e2e = None
    if not args.no_e2e:
        # pinned staging (as the contract asks) unless the set is so large that pinning it on every rank
        # of the box would lock > 64 GB of host memory (8 ranks x 23 GB at N = 8): then pageable
        pin = desc_full.numel() <= 8 * 2**30
        h_desc = torch.empty(desc_full.shape, dtype=torch.uint8, pin_memory=pin)
        h_desc.copy_(desc_full)
        h_np = h_desc.numpy().reshape(n_img, K, 128)
        descs = [h_np[i] for i in range(n_img)]
        kp = None
        if verify:
            h_k = torch.empty(kpts_full.shape, dtype=torch.float32, pin_memory=pin)
            h_k.copy_(kpts_full)
            kp = [h_k.numpy().reshape(n_img, K, 2)[i] for i in range(n_img)]

        def e2e_step():
            ctx.set_images(descs, kp, cams if verify else None)
            res = ctx.match_pairs(my_pairs, sift, tvg)
            nm = res.total_matches
            n_in = 0
            v = res.view(len(my_pairs) - 1)      # touch the result object like a caller would
            _ = v.n_matches
            res.free()
            return nm, n_in
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            nm, _ = e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = n_img * K * 128 + (n_img * K * 8 if verify else 0) + len(my_pairs) * 8
        d2h = nm * 8 + len(my_pairs) * 12
        e2e = {"value": pairs_total / te.item(), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h),
               "note": "b2m_set_images(host) + b2m_match_pairs + results in host memory, wall clock, max over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1: int8 GEMM + fused top-2), tensor-bound
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops_sustained")
    peak_src = "2 x MEASURED_PEAKS.json bf16_tflops_sustained (derived int8 peak: kind::i8 runs at twice the bf16 rate)"
    if not bf16:
        bf16, peak_src = 1400.0, "2 x fallback sustained bf16 1.4 PFLOP/s (B200_PROFILING.md), derived int8 peak"
    peak = 2.0 * bf16
    try:
        # measured on this pool's B200 by tools/microbench.cu (tcgen05.mma.kind::i8 issue loop, all SMs)
        mb = json.load(open(os.path.join(ROOT, "profiles", "r01_microbench_tmem_i8mma.json")))
        peak = float(mb["i8_mma_n256_chip_TOPS"])
        peak_src = ("measured int8 tcgen05 peak, tools/microbench.cu on this pool's B200 "
                    "(profiles/r01_microbench_tmem_i8mma.json, burst, 1965 MHz)")
    except Exception:
        pass
    ops_per_pair = 2.0 * K * K * 128
    k1_avg_ms = k1_ms / max(k1_n, 1)
    pairs_per_launch = len(my_pairs) * args.steps / max(k1_n, 1)
    achieved = ops_per_pair * pairs_per_launch / (k1_avg_ms / 1e3) / 1e12
    # 1 / 4: the column direction of the cross-check runs only for pairs with row-direction candidates
    # (two launches of the GEMM kernel per batch, include/b200match.h enum b2m_k1_dir1_mode); else one launch
    dir1_mode = int(ctx.stats().k1_dir1_mode)
    split = dir1_mode in (1, 4)
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch (1024 pairs of 8192^2) from the
            # DRAM bytes of one 1024-pair launch, ncu --set full capture profiles/r01_k1_v14_final.ncu_summary.txt:
            # 136.8 MB read + 59.3 MB written (the distance matrix never leaves TMEM; images mostly hit in L2)
            # (that capture is of the single two-direction launch; the split schedule has no capture yet -> null)
            "traffic": 196.1e6 if (K == 8192 and not split) else None,
            "kernel": ("b2m_k1_filter_kernel x2 per batch (all pairs row direction + live pairs column direction)"
                       if split else "b2m_k1_filter_kernel"),
            "k1_dir1_mode": dir1_mode, "avg_launch_ms": k1_avg_ms,
            "pairs_per_launch": pairs_per_launch, "peak_source": peak_src,
            "algorithmic": "2*K1*K2*128 int8 ops per pair (one GEMM; the transposed GEMM of the cross-check "
                           "direction is not counted)"}

    if rank == 0:
        log("e2e leg done; cpu baseline ...")
    cb = None
    if not args.no_cpu:
        n_small = min(n_img, 48)
        cb, sample, cpu_res = cpu_baseline(desc_full[: n_small * K].cpu().numpy(), np.full(n_small, K, np.int32),
                                           syn.exhaustive_pairs(n_small), args.cpu_seconds, verify,
                                           kpts_full[: n_small * K].cpu().numpy(), cams[0],
                                           full_frac=verified_fraction if verify else None)
        # the same sample through the GPU path must be bit-identical
        chk = ctx.match_pairs(sample, sift, None)
        same = all(np.array_equal(chk.matches(k), cpu_res[k]) for k in range(len(sample)))
        cb["gpu_bit_exact_on_sample"] = bool(same)
        chk.free()

    out = {
        "metric": metric, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "images": n_img, "features_per_image": K, "pairs_per_step": int(pairs_total),
                   "matches_per_step_rank0": int(total_matches), "verified_pairs_fraction_rank0": verified_fraction, "parallelism": f"pair-sharded x{world}",
                   "l2": "inputs (descriptor set %.2f GB) larger than L2" % (n_img * K * 128 / 1e9),
                   "timing": "CUDA events on the library stream around each b2m_match_pairs call, max over ranks"},
        "wall_ms_per_step": wall_ms_max / args.steps, "k1_ms_per_step": k1_ms / args.steps,
        "compact_verify_ms_per_step": ver_ms / args.steps, "gpu_launches": int(launches_total), "clocks": clk,
        "roofline": roof, "cpu_baseline": cb, "e2e": e2e,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
