# -*- coding: utf-8 -*-
"""
Node sharding of the coalescence scan across GPUs.

Nodes are independent in ``migrate`` (``quakemigrate/core/src/migratelib.c:51-64``); only the
per-sample reduction of ``find_max_coa`` (``migratelib.c:99-109``) crosses nodes.  So:

* rank r owns a contiguous slab of the C-order flat node list -- whole x-planes
  ``[x0, x1)`` of the ``(nx, ny, nz)`` grid (:func:`slab_bounds`), uploads only that part of the
  travel-time table together with ``node_offset = x0 * ny * nz`` (:func:`shard_traveltimes`);
* every rank receives identical onset functions;
* per detect call the ranks exchange, per output sample, their local top candidates and node-sum
  (one NCCL all-gather inside ``libqmb200``), then every rank folds them identically.

:func:`merge_rank_results` is the host-side statement of what that fold must produce from
per-rank *final* results (max, normalised max, arg-max, local node count); the world-size-2
``gloo`` tests use it to pin the sharding semantics on CPU.
"""

from __future__ import annotations

import numpy as np


def slab_bounds(nx: int, world: int, rank: int) -> tuple[int, int]:
    """x-plane range [x0, x1) of ``rank``: planes are dealt as evenly as possible, in order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {rank}/{world}")
    if nx < world:
        raise ValueError(f"cannot shard {nx} x-planes over {world} ranks")
    base, extra = divmod(nx, world)
    x0 = rank * base + min(rank, extra)
    return x0, x0 + base + (1 if rank < extra else 0)


def balanced_range(nx: int, ny: int, nz: int, world: int, rank: int) -> tuple[int, int]:
    """
    ``(node_first, n_nodes_local)`` of ``rank`` when the C-order node list is cut into ``world``
    contiguous ranges of (almost) equal size in whole z-columns -- the unit the device tiling needs.
    Unlike :func:`slab_bounds` the cuts may fall inside an x-plane: 201 x-planes over 8 ranks are 26
    or 25 planes each (3.5 % imbalance), 40401 z-columns are 5051 or 5050 (0.02 %).
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {rank}/{world}")
    columns = nx * ny
    if columns < world:
        raise ValueError(f"cannot shard {columns} z-columns over {world} ranks")
    base, extra = divmod(columns, world)
    c0 = rank * base + min(rank, extra)
    return c0 * nz, (base + (1 if rank < extra else 0)) * nz


def shard_traveltimes_balanced(traveltimes: np.ndarray, world: int, rank: int):
    """Rows ``(n_local, S)`` of ``rank``'s balanced range; returns (rows, node_first, n_nodes_global)."""
    nx, ny, nz, s = traveltimes.shape
    first, count = balanced_range(nx, ny, nz, world, rank)
    flat = traveltimes.reshape(-1, s)
    return np.ascontiguousarray(flat[first:first + count]), first, nx * ny * nz


def shard_traveltimes(traveltimes: np.ndarray, world: int, rank: int):
    """Slice ``(nx, ny, nz, S)`` for ``rank``; returns (slab, node_offset, n_nodes_global)."""
    nx, ny, nz, _ = traveltimes.shape
    x0, x1 = slab_bounds(nx, world, rank)
    return np.ascontiguousarray(traveltimes[x0:x1]), x0 * ny * nz, nx * ny * nz


def merge_rank_results(per_rank):
    """
    Fold per-rank detect results into the global one.

    ``per_rank``: list (in rank order) of ``(max_coa, max_norm_coa, max_idx_global, n_nodes_local)``
    where each rank's normalised max is ``max * n_local / sum_local`` (migratelib.c:108 applied to
    its slab).  Global: the largest max wins, ties go to the LOWEST global node index
    (migratelib.c:96-105 visits nodes in ascending order with a strict ``>``); the node sums add.
    """
    n_total = sum(int(r[3]) for r in per_rank)
    best = np.array(per_rank[0][0], dtype=np.double, copy=True)
    idx = np.array(per_rank[0][2], dtype=np.int64, copy=True)
    total = per_rank[0][0] * per_rank[0][3] / per_rank[0][1]
    for mx, nm, ix, n_local in per_rank[1:]:
        total = total + mx * n_local / nm
        take = (mx > best) | ((mx == best) & (ix < idx))
        best = np.where(take, mx, best)
        idx = np.where(take, ix, idx)
    return best, best * n_total / total, idx
