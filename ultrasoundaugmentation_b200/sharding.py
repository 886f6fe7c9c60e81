"""Multi-GPU sharding: every image is an independent unit, so a batch splits across ranks with
NO collective on the data path.  Parameters are keyed by the GLOBAL sample
index, so every rank draws exactly the parameters the 1-rank run draws for its samples: parameters, labels and
the K1 / K3 images of the N-rank run equal the 1-rank run bit for bit.  Confidence maps (and the chain images
that depend on them) are bitwise equal when both runs select the same solver variant -- the library picks it
from the size of a launch (csrc/confmap_pcg.cu: cm_layout) -- and agree within the confidence-map tolerance
otherwise; pin the variant (ops.confidence_map: coarse / twist / up_variant) for bitwise results at any N.
An optional final gather (NCCL on GPUs, gloo in CPU tests) is provided for consumers that want the whole batch.
"""
from __future__ import annotations

import torch


def shard_bounds(total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous split; the first (total % world_size) ranks get one extra item."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    q, r = divmod(int(total), int(world_size))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_indices(total: int, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(total, rank, world_size)
    return torch.arange(lo, hi, dtype=torch.int64)


def shard_params(chain, total: int, rank: int, world_size: int, image_size, generator=None) -> dict:
    """Parameters of this rank's shard, identical to rows [lo:hi) of the single-rank draw."""
    idx = shard_indices(total, rank, world_size).numpy()
    return chain.get_params(len(idx), generator, idx, image_size)


def gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Optional final all-gather of per-rank shards [n_r, ...] into [total, ...] on every rank.

    Not on the hot path.  Ragged shards are padded to the largest shard for the collective.
    """
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    sizes = [shard_bounds(total, r, ws) for r in range(ws)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < nmax:
        pad = torch.cat([local, local.new_zeros((nmax - local.shape[0],) + tuple(local.shape[1:]))])
    out = local.new_empty((ws * nmax,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * nmax: r * nmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts)
