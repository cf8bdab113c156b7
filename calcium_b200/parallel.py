"""Sharding an ensemble over the GPUs of one box.

Pouches are independent (the reference runs them one after another, ``main.py:202``), so the
multi-GPU path is a pure partition: no data-path collective, no NCCL.  Each rank integrates and
rasterises its own pouches; ``torch.distributed`` is only used for the timing barrier / max in the
benchmark and, optionally, to gather small per-pouch summaries on rank 0.

Partition: longest-processing-time-first over the cost model ``n_cells * (T-1)`` (K1 dominates); for
equal-size ensembles this degenerates to an even split.  The assignment is a pure function of the
pouch list and the world size, so every rank computes it locally and they all agree.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def pouch_cost(n_cells: int, T: int, n_frames: int = 0, pixels: int = 0) -> float:
    """Relative cost: integration (FP64/shared-memory bound) + frame writes (HBM bound).
    The second term is scaled so that one 512x512 frame costs about as much as 350 cell-steps
    (measured ratio of K2 to K1 time on B200)."""
    return float(n_cells) * max(T - 1, 0) + 350.0 * n_frames * (pixels / 262144.0)


def lpt_partition(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-first assignment; returns, per rank, the sorted list of pouch indices."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    count = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], count[k], k))
        parts[r].append(i)
        load[r] += costs[i]
        count[r] += 1
    return [sorted(p) for p in parts]


def shard_indices(n_cells: Sequence[int], T: int, world_size: int, n_frames: int = 0,
                  pixels: int = 0) -> List[List[int]]:
    return lpt_partition([pouch_cost(n, T, n_frames, pixels) for n in n_cells], world_size)


def frames_digest(frames: np.ndarray) -> np.ndarray:
    """Order-independent-free 2x uint64 checksum of a frames block (bit pattern sum and xor);
    used to compare sharded and unsharded runs without moving the frames."""
    u = np.ascontiguousarray(frames).view(np.uint64).ravel()
    return np.array([np.bitwise_xor.reduce(u) if u.size else 0,
                     np.add.reduce(u, dtype=np.uint64) if u.size else 0], dtype=np.uint64)
