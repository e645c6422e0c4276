"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed.

Two modes, both prescribed by BASELINE.json north_star:

* scenario sharding -- independent scenarios (tau_aer x mu0 x albedo x phase sweeps) are dealt to the
  ranks; no data-path communication, one gather of the small per-scenario results at the end.
* mu-block sharding -- one very large grid (config 4: ~10k layers x 1024 mu): rank g owns a block of mu
  columns.  Each order it computes J[:, block] = I_{n-1}[:, all] . A[all, block] (source contraction
  restricted to its column tiles), sweeps its own columns, and the ranks ALL-GATHER their I_n blocks so
  that every rank holds the full I_n for the next contraction; the two convergence ratios are reduced
  with a MAX all-reduce.  NCCL over NVLink carries both collectives.

The helpers below are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

TILE = 128  # column granularity of the source contraction


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) share of n_items for `rank`."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_scenarios(scenarios: Sequence, rank: int, world: int) -> List:
    """Round-robin deal (keeps expensive neighbours of a sorted sweep on different ranks)."""
    return [sc for i, sc in enumerate(scenarios) if i % world == rank]


def gather_scenario_results(local: List, n_total: int, rank: int, world: int, group=None) -> Optional[List]:
    """Inverse of shard_scenarios: every rank contributes its small result objects; rank 0 gets the list."""
    if world == 1:
        return list(local)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * n_total
    for r, part in enumerate(bucket):
        for j, item in enumerate(part):
            out[j * world + r] = item
    return out


def mu_blocks(N: int, M: int, world: int, zone_lo: int) -> List[Tuple[int, int]]:
    """Column blocks [c0, c1) per rank: multiples of TILE, balanced in tiles, and never cutting the
    mu -> 0 zones: columns [zone_lo, M) (extrapolation sources/targets, windowed columns) stay in one
    block and the boundary M+1 (start of the upward blend zone) is never a cut."""
    ntiles = (N + TILE - 1) // TILE
    if world > ntiles:
        raise ValueError(f"{world} ranks for {ntiles} column tiles")
    cuts = [0]
    for r in range(1, world):
        cuts.append(shard_range(ntiles, r, world)[0] * TILE)
    cuts.append(N)
    for c in cuts[1:-1]:
        if zone_lo < c < M or c == M + 1:
            raise ValueError(f"block boundary {c} cuts the mu->0 zone [{zone_lo}, {M}) / {M + 1}")
        if M + 1 < c < min(N, M + 1 + 64):
            raise ValueError(f"block boundary {c} leaves fewer than 64 columns for the upward blend search")
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def allgather_columns(field: torch.Tensor, blocks: Sequence[Tuple[int, int]], rank: int, group=None) -> None:
    """In place: every rank contributes field[:, c0:c1] of its own block and receives all others."""
    world = len(blocks)
    if world == 1:
        return
    width = max(c1 - c0 for c0, c1 in blocks)
    rows = field.shape[0]
    send = torch.zeros((rows, width), dtype=field.dtype, device=field.device)
    c0, c1 = blocks[rank]
    send[:, : c1 - c0].copy_(field[:, c0:c1])
    recv = torch.empty((world, rows, width), dtype=field.dtype, device=field.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(recv, send, group=group)   # one ncclAllGather, no staging copies
    else:
        dist.all_gather([recv[r] for r in range(world)], send, group=group)
    for r, (a, b) in enumerate(blocks):
        if r != rank:
            field[:, a:b].copy_(recv[r, :, : b - a])


class MuShardedSolver:
    """Order loop of one large single-layer grid sharded by mu blocks (config 4)."""

    def __init__(self, engine, blocks, rank, group=None):
        self.eng, self.blocks, self.rank, self.group = engine, list(blocks), rank, group
        self.world = len(self.blocks)
        c0, c1 = self.blocks[rank]
        engine.set_columns(c0, c1)
        self.comm_ms = 0.0

    def solve(self, I1: torch.Tensor, max_orders: int = 10000, poll_every: int = 4, time_comm: bool = False):
        eng = self.eng
        I = I1.clone()
        In = I1.clone()
        J = eng.new_field(zero=True)
        eng.reset(I1)
        ratios = torch.empty((eng.S, 2), dtype=torch.float64, device=eng.device)
        n = 1
        done = False
        ev = []
        while n < max_orders and not done:
            n += 1
            eng.source(In, out=J)
            eng.sweeps(J, out=In, accumulate_into=I)
            if time_comm:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            allgather_columns(In, self.blocks, self.rank, self.group)
            if self.world > 1:
                eng.ratios(ratios)
                dist.all_reduce(ratios, op=dist.ReduceOp.MAX, group=self.group)
                eng.ratios(ratios, set=True)
            if time_comm:
                e1.record()
                ev.append((e0, e1))
            eng.converge(n)
            if (n - 1) % poll_every == 0:
                done = not any(r.active for r in eng.results())
        res = eng.results()
        allgather_columns(I, self.blocks, self.rank, self.group)
        if time_comm:
            torch.cuda.synchronize()
            self.comm_ms = sum(a.elapsed_time(b) for a, b in ev)
        return I, res
