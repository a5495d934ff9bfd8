"""Chunk-position lists of the BASELINE.json configs and slab sharding across GPUs.

Enumeration order is x outer, y, z inner -- the order `World::update_nearby` walks its window
(world.rs:164-170).  Chunks are independent (a chunk's output depends only on pos, seed and
constants: chunk.rs:89-129), so a region is sharded as contiguous x-slabs, one per rank, with
no collective on the compute path.  Never cut along z: only chunk layers z in [-3, 1] can hold
surface, so z-slabs would not balance.
"""
from __future__ import annotations

import numpy as np


def box_region(x_range, y_range, z_range) -> np.ndarray:
    """All chunk positions in [x0,x1) x [y0,y1) x [z0,z1), x-major, z fastest.  int32 [n,3]."""
    xs = np.arange(x_range[0], x_range[1], dtype=np.int32)
    ys = np.arange(y_range[0], y_range[1], dtype=np.int32)
    zs = np.arange(z_range[0], z_range[1], dtype=np.int32)
    g = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), axis=-1)
    return np.ascontiguousarray(g.reshape(-1, 3))


def config_positions(name: str) -> np.ndarray:
    """BASELINE.md §5 inputs."""
    if name == "single":            # config 1
        return np.array([[0, 0, 0]], dtype=np.int32)
    if name == "spawn":             # config 2: 16x16x8 = 2048 chunks around the sub start chunk (0,0,0)
        return box_region((-8, 8), (-8, 8), (-4, 4))
    if name == "large":             # config 3: 128x128x32 = 524,288 chunks
        return box_region((-64, 64), (-64, 64), (-16, 16))
    raise ValueError(name)


def slab_bounds(n_x: int, rank: int, world: int) -> tuple:
    """Contiguous x-slab [lo, hi) of rank `rank`; the remainder is spread one-per-rank."""
    base, rem = divmod(n_x, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_region(x_range, y_range, z_range, rank: int, world: int) -> np.ndarray:
    """Positions of rank `rank`'s x-slab of the box (strong scaling: the box is fixed)."""
    lo, hi = slab_bounds(x_range[1] - x_range[0], rank, world)
    return box_region((x_range[0] + lo, x_range[0] + hi), y_range, z_range)


def weak_region(x_per_rank: int, y_range, z_range, rank: int) -> np.ndarray:
    """Rank `rank`'s slab when every rank owns `x_per_rank` x-columns (weak scaling): the region
    grows along x with the number of ranks, centred like config 3."""
    x0 = rank * x_per_rank
    return box_region((x0, x0 + x_per_rank), y_range, z_range)
