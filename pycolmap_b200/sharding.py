"""Multi-GPU plumbing: image pairs are independent units.

One process per GPU (`torch.distributed`, NCCL on GPUs / gloo in the CPU tests).  Every rank
owns a contiguous shard of images; ONE all-gather makes the whole descriptor set resident on every
rank, pairs are dealt round-robin, and there is no further data-path collective.  Matching results
do not depend on batching, so they are bit-identical for any world size.
"""
import numpy as np


def image_shard(n_images, rank, world):
    """Contiguous image range [lo, hi) of `rank`; shards are ceil(n/world) wide (last ones may be short)."""
    per = (n_images + world - 1) // world
    return min(rank * per, n_images), min((rank + 1) * per, n_images), per


def pair_shard(pairs, rank, world):
    """Round-robin deal of the pair list: equal counts (+-1) and, for exhaustive lists ordered by
    first image, an even mix of near and far pairs on every rank."""
    return np.ascontiguousarray(np.asarray(pairs)[rank::world])


def all_gather_rows(local, n_images, rows_per_image, rank, world, dist=None):
    """All-gather per-image row blocks ([n_local * rows_per_image, C] tensors) into the full
    [n_images * rows_per_image, C] tensor on every rank (one collective)."""
    import torch
    if world == 1:
        return local
    lo, hi, per = image_shard(n_images, rank, world)
    pad = per - (hi - lo)
    if pad:
        local = torch.cat([local, torch.zeros((pad * rows_per_image,) + tuple(local.shape[1:]), dtype=local.dtype,
                                              device=local.device)])
    full = torch.empty((world * per * rows_per_image,) + tuple(local.shape[1:]), dtype=local.dtype,
                       device=local.device)
    dist.all_gather_into_tensor(full, local.contiguous())
    return full[: n_images * rows_per_image]
