"""Chain sharding over ranks and the final gather of per-chain minima.

Chains are independent, so the annealing itself needs no collective: rank r owns the contiguous global
chain range chain_range(r, world, n_total) and seeds / parameter sets are indexed by the GLOBAL chain id,
which makes every per-chain result independent of the number of GPUs.  The only exchange is the all-gather
of the packed tr_minimum records (24 bytes per chain) at the end, after which every rank knows the global
winner and its owner fetches the structure (what Space.writeMinEnergy needs, Space.java:836-844).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .engine import MINIMUM_DTYPE


def chain_range(rank: int, world: int, n_total: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns chains [lo, hi); the first n_total % world ranks get one more."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(chain: int, world: int, n_total: int) -> int:
    for r in range(world):
        lo, hi = chain_range(r, world, n_total)
        if lo <= chain < hi:
            return r
    raise ValueError(f"chain {chain} outside 0..{n_total - 1}")


def gather_minima(local, world: int, counts=None):
    """All-gather packed tr_minimum records over torch.distributed (NCCL for CUDA tensors, gloo for CPU).

    local: uint8 torch tensor holding this rank's MINIMUM_DTYPE records (device or host).  Returns a numpy
    structured array of every rank's records, ordered by global chain id.  With world == 1 no collective
    is issued."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local.cpu().numpy().view(MINIMUM_DTYPE).copy()
    n_local = local.numel()
    if counts is None:
        out = torch.empty(world * n_local, dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(out, local)
        parts = [out]
    else:  # ragged shards: pad to the largest, trim afterwards
        width = max(counts) * MINIMUM_DTYPE.itemsize
        padded = torch.zeros(width, dtype=torch.uint8, device=local.device)
        padded[:n_local] = local
        out = torch.empty(world * width, dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(out, padded)
        parts = [out[r * width : r * width + counts[r] * MINIMUM_DTYPE.itemsize] for r in range(world)]
    arr = np.concatenate([p.cpu().numpy().view(MINIMUM_DTYPE) for p in parts])
    return arr[np.argsort(arr["chain"], kind="stable")]


def global_minimum(minima: np.ndarray) -> Tuple[float, int, int]:
    """(energy, global chain id, tooth) of the lowest write(n) snapshot over all chains; ties go to the lowest
    chain id, like the strict `<` of Space.java:614 goes to the earliest snapshot."""
    i = int(np.argmin(minima["energy"]))
    return float(minima["energy"][i]), int(minima["chain"][i]), int(minima["tooth"][i])
