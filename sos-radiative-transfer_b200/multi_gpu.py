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


class ColumnGather:
    """Preallocated buffers for repeated column all-gathers of equally shaped row slabs (keeps the
    per-call host work to three tensor ops: pack, ncclAllGather, unpack)."""

    def __init__(self, rows: int, blocks: Sequence[Tuple[int, int]], rank: int, dtype, device, group=None):
        self.blocks, self.rank, self.group = list(blocks), rank, group
        self.world = len(self.blocks)
        self.width = max(c1 - c0 for c0, c1 in self.blocks)
        self.uniform = all(c1 - c0 == self.width for c0, c1 in self.blocks) and \
            all(self.blocks[r][0] == r * self.width for r in range(self.world))
        self.send = torch.zeros((rows, self.width), dtype=dtype, device=device)
        self.recv = torch.empty((self.world, rows, self.width), dtype=dtype, device=device)
        self.nccl = self.world > 1 and dist.get_backend(group) == "nccl"

    def __call__(self, field: torch.Tensor) -> None:
        if self.world == 1:
            return
        rows = field.shape[0]
        c0, c1 = self.blocks[self.rank]
        send, recv = self.send[:rows], self.recv[:, :rows]
        send[:, : c1 - c0].copy_(field[:, c0:c1])
        if rows != self.send.shape[0]:
            send, recv = send.contiguous(), torch.empty((self.world, rows, self.width), dtype=field.dtype, device=field.device)
        if self.nccl:
            dist.all_gather_into_tensor(recv, send, group=self.group)   # one ncclAllGather, no staging copies
        else:
            dist.all_gather([recv[r] for r in range(self.world)], send, group=self.group)
        if self.uniform:
            field[:, : self.world * self.width].view(rows, self.world, self.width).copy_(recv.permute(1, 0, 2))
        else:
            for r, (a, b) in enumerate(self.blocks):
                if r != self.rank:
                    field[:, a:b].copy_(recv[r, :, : b - a])


def allgather_columns(field: torch.Tensor, blocks: Sequence[Tuple[int, int]], rank: int, group=None) -> None:
    """In place: every rank contributes field[:, c0:c1] of its own block and receives all others."""
    if len(blocks) == 1:
        return
    ColumnGather(field.shape[0], blocks, rank, field.dtype, field.device, group)(field)


class RawField:
    """A device buffer that is not a torch tensor (CUDA-IPC shareable allocation); quacks like one for
    the engine calls, which only need data_ptr()."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = int(ptr), int(nbytes)

    def data_ptr(self) -> int:
        return self.ptr


class PeerFields:
    """Two ping-pong I_n buffers per rank, mapped into every rank's address space with CUDA IPC, so that
    the source contraction can read each mu block from the GPU that owns it (sos_source_peers)."""

    def __init__(self, engine, rank: int, world: int, group=None):
        import ctypes as C
        from . import _lib
        self.lib, self.C, self._lib = _lib.load(), C, _lib
        self.rank, self.world = rank, world
        nbytes = engine.S * engine.L * engine.ld * 8
        self.nbytes = nbytes
        self.local, handles = [], []
        with torch.cuda.device(engine.device):
            for _ in range(2):
                ptr = C.c_void_p()
                h = C.create_string_buffer(64)
                _lib.check(self.lib.sos_ipc_alloc(nbytes, C.byref(ptr), h), "sos_ipc_alloc")
                self.local.append(RawField(ptr.value, nbytes))
                handles.append(h.raw)
            everyone = [None] * world
            dist.all_gather_object(everyone, handles, group=group)
            self.ptrs = []          # [buffer][rank] -> device pointer valid in THIS process
            self._opened = []
            for b in range(2):
                row = []
                for r in range(world):
                    if r == rank:
                        row.append(self.local[b].ptr)
                    else:
                        ptr = C.c_void_p()
                        _lib.check(self.lib.sos_ipc_open(everyone[r][b], C.byref(ptr)), "sos_ipc_open")
                        self._opened.append(ptr.value)
                        row.append(ptr.value)
                self.ptrs.append(row)

    def ptr_array(self, b: int):
        C = self.C
        return (C.c_void_p * self.world)(*self.ptrs[b])

    def close(self):
        for p in self._opened:
            self.lib.sos_ipc_close(self.C.c_void_p(p))
        self._opened = []
        for f in self.local:
            self.lib.sos_ipc_free(self.C.c_void_p(f.ptr))
        self.local = []


class PeerBuffers:
    """Device buffers of the given sizes on every rank, each mapped into every rank's address space with CUDA IPC.
    ptrs[b][r] = buffer b of rank r as a device pointer valid in THIS process."""

    def __init__(self, sizes: Sequence[int], device, rank: int, world: int, group=None):
        import ctypes as C
        from . import _lib
        self.lib, self.C = _lib.load(), C
        self.rank, self.world, self.sizes = rank, world, [int(n) for n in sizes]
        self.local, handles, self._opened = [], [], []
        with torch.cuda.device(device):
            for nbytes in self.sizes:
                ptr = C.c_void_p()
                h = C.create_string_buffer(64)
                _lib.check(self.lib.sos_ipc_alloc(nbytes, C.byref(ptr), h), "sos_ipc_alloc")   # (zero-filled)
                self.local.append(RawField(ptr.value, nbytes))
                handles.append(h.raw)
            everyone = [handles]
            if world > 1:
                everyone = [None] * world
                dist.all_gather_object(everyone, handles, group=group)
            self.ptrs = []
            for b in range(len(self.sizes)):
                row = []
                for r in range(world):
                    if r == rank:
                        row.append(self.local[b].ptr)
                    else:
                        ptr = C.c_void_p()
                        _lib.check(self.lib.sos_ipc_open(everyone[r][b], C.byref(ptr)), "sos_ipc_open")
                        self._opened.append(ptr.value)
                        row.append(ptr.value)
                self.ptrs.append(row)

    def ptr_array(self, b: int):
        return (self.C.c_void_p * self.world)(*self.ptrs[b])

    def close(self):
        for p in self._opened:
            self.lib.sos_ipc_close(self.C.c_void_p(p))
        self._opened = []
        for f in self.local:
            self.lib.sos_ipc_free(self.C.c_void_p(f.ptr))
        self.local = []


class LayerShardedSolver:
    """Order loop of one large single-layer grid sharded by LAYER blocks (BASELINE configs[3]).

    Rank r owns a contiguous block of scan chunks.  The source contraction is row-local (no communication); per order the
    ranks exchange the chunk aggregates of the scan and the halo rows next to a block boundary by stores into each other's
    memory over NVLink, issued from kernels inside the CUDA-graphed order loop of sos_solve (csrc/layer_shard.cuh) -- no NCCL
    call and no host round trip per order.  The result is bit-identical to the unsharded solve of a plan with the same scan
    chunks.  NCCL carries only the rendezvous and the final gather of the row blocks of I."""

    def __init__(self, engine, rank: int, world: int, group=None, emulate=None):
        """emulate=(rank, world): ONE process plays one rank of a larger job with its own buffers standing in for the peers'
        (profiling a rank's kernels under ncu with SOS_B200_PEER_TIMEOUT_MS=0; the numbers it computes are meaningless)."""
        self.eng, self.rank, self.world, self.group = engine, rank, world, group
        nfield = engine.S * engine.L * engine.ld * 8
        self.bufs = PeerBuffers([engine.layer_mailbox_bytes(), nfield], engine.device, rank, world, group)
        self.In = self.bufs.local[1]
        if emulate is not None:
            import ctypes as C
            er, ew = emulate
            row0, row1 = engine.set_layers(er, ew, (C.c_void_p * ew)(*[self.bufs.local[0].ptr] * ew), (C.c_void_p * ew)(*[self.In.ptr] * ew))
        else:
            row0, row1 = engine.set_layers(rank, world, self.bufs.ptr_array(0), self.bufs.ptr_array(1))
        rows = [(row0, row1)]
        if world > 1:
            rows = [None] * world
            dist.all_gather_object(rows, (row0, row1), group=group)
        self.rows = rows
        self._I = None
        self._stage = None
        if world > 1:
            torch.cuda.synchronize(engine.device)
            dist.barrier(group=group)      # every rank's mailbox is zeroed and mapped before anybody writes into it

    def solve(self, I1: torch.Tensor, max_orders: int = 10000, gather: bool = True):
        """I1: the first order on ALL rows (every rank computes it: closed form, SOS_Aer_I1_In.py:13-58).  Returns (I, results);
        with gather=True every rank ends with the full field, otherwise only its own rows of I are final."""
        import ctypes as C
        from . import _lib
        eng = self.eng
        if self._I is None:
            self._I = torch.empty_like(I1)
            self._J = eng.new_field(zero=True)
        I, J = self._I, self._J
        I.copy_(I1)
        res = (_lib.sos_result * eng.S)()
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.sos_copy_d2d(C.c_void_p(self.In.ptr), C.c_void_p(I1.data_ptr()), self.In.nbytes, eng._stream), "sos_copy_d2d")
            _lib.check(eng.lib.sos_solve(eng._plan, I.data_ptr(), C.c_void_p(self.In.ptr), J.data_ptr(), None, 0, int(max_orders), 1, res,
                                         eng._stream), "sos_solve")
        if gather and self.world > 1:
            self.gather_rows(I)
        return I, res

    def gather_rows(self, field: torch.Tensor) -> None:
        """In place: every rank contributes its own rows of `field` and receives everybody else's."""
        L, ld = self.eng.L, self.eng.ld
        width = max(b - a for a, b in self.rows)
        if self._stage is None:
            self._stage = (torch.zeros((width, ld), dtype=field.dtype, device=field.device),
                           torch.empty((self.world, width, ld), dtype=field.dtype, device=field.device))
        send, recv = self._stage
        a, b = self.rows[self.rank]
        f2 = field.view(-1, ld)[:L]
        send[: b - a].copy_(f2[a:b])
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(recv, send, group=self.group)
        else:   # gloo (CPU tests, or several ranks sharing one GPU): through host memory
            parts = [torch.empty((width, ld), dtype=field.dtype) for _ in range(self.world)]
            dist.all_gather(parts, send.cpu(), group=self.group)
            recv.copy_(torch.stack(parts))
        for r, (a, b) in enumerate(self.rows):
            if r != self.rank:
                f2[a:b].copy_(recv[r, : b - a])

    def close(self):
        torch.cuda.synchronize(self.eng.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        self.eng.set_layers(0, 1)
        self.bufs.close()


class MuShardedSolver:
    """Order loop of one large single-layer grid sharded by mu blocks (config 4)."""

    def __init__(self, engine, blocks, rank, group=None):
        self.eng, self.blocks, self.rank, self.group = engine, list(blocks), rank, group
        self.world = len(self.blocks)
        c0, c1 = self.blocks[rank]
        engine.set_columns(c0, c1)
        self.comm_ms = 0.0

    def solve_p2p(self, I1: torch.Tensor, peers: "PeerFields", max_orders: int = 10000, poll_every: int = 8):
        """Order loop with the all-gather FUSED into the contraction: every rank keeps its I_n block in
        a CUDA-IPC buffer and the contraction kernel of order n+1 pulls each k-range from its owner by TMA
        over NVLink (sos_source_peers).  The only collective left per order is the MAX all-reduce of the
        two convergence ratios, which doubles as the cross-GPU barrier that orders writers and readers of
        the ping-pong buffers."""
        import ctypes as C
        from . import _lib
        eng, lib = self.eng, self.eng.lib
        dev = eng.device
        if getattr(self, "_p2p_bufs", None) is None:   # work buffers are kept between solves
            self._p2p_bufs = (torch.empty_like(I1), eng.new_field(zero=True),
                              torch.empty((eng.S, 2), dtype=torch.float64, device=dev))
        I, J, ratios = self._p2p_bufs
        I.copy_(I1)
        eng.reset(I1)
        cols = (C.c_int * (self.world + 1))(*([b[0] for b in self.blocks] + [self.blocks[-1][1]]))
        cur = 0
        with torch.cuda.device(dev):
            _lib.check(lib.sos_copy_d2d(C.c_void_p(peers.local[cur].ptr), C.c_void_p(I1.data_ptr()), peers.nbytes, eng._stream), "sos_copy_d2d")
        torch.cuda.current_stream(dev).synchronize()
        dist.barrier(group=self.group)             # every rank's I1 block is in place

        def one_order(b, order):
            with torch.cuda.device(dev):
                _lib.check(lib.sos_source_peers(eng._plan, peers.ptr_array(b), self.world, cols, J.data_ptr(), eng._stream),
                           "sos_source_peers")
            eng.sweeps(J, out=peers.local[b ^ 1], accumulate_into=I)
            eng.ratios(ratios)
            dist.all_reduce(ratios, op=dist.ReduceOp.MAX, group=self.group)
            eng.ratios(ratios, set=True)
            eng.converge(order)

        n = 1
        done = False
        while n < max_orders and not done:
            n += 1
            one_order(cur, n)
            cur ^= 1
            if (n - 1) % poll_every == 0:
                done = not any(r.active for r in eng.results())
        res = eng.results()
        allgather_columns(I, self.blocks, self.rank, self.group)
        return I, res

    def solve(self, I1: torch.Tensor, max_orders: int = 10000, poll_every: int = 4, time_comm: bool = False,
              row_chunks: int = 4):
        """Order loop.  After the sweeps of order n the I_n blocks are all-gathered in `row_chunks` slabs of
        layers on a side stream; the contraction of order n+1 is launched slab by slab as soon as its
        slab has arrived, so NVLink traffic overlaps the FP64 work."""
        eng = self.eng
        dev = eng.device
        I = I1.clone()
        In = I1.clone()
        J = eng.new_field(zero=True)
        eng.reset(I1)
        ratios = torch.empty((eng.S, 2), dtype=torch.float64, device=dev)
        L = eng.L
        step = max(64, -(-L // max(row_chunks, 1) // 64) * 64) if self.world > 1 else L
        slabs = [(r, min(L, r + step)) for r in range(0, L, step)]
        gather = ColumnGather(step, self.blocks, self.rank, In.dtype, dev, self.group)
        compute = torch.cuda.current_stream(dev)
        comm = torch.cuda.Stream(device=dev) if self.world > 1 else compute
        n = 1
        done = False
        ev = []
        eng.source(In, out=J)                      # order 2 needs no exchange: every rank holds the full I1
        while n < max_orders and not done:
            n += 1
            eng.sweeps(J, out=In, accumulate_into=I)
            if self.world == 1:
                eng.converge(n)
                eng.source(In, out=J)
            else:
                swept = torch.cuda.Event(enable_timing=time_comm)
                swept.record(compute)
                comm.wait_event(swept)
                with torch.cuda.stream(comm):
                    eng.ratios(ratios)
                    dist.all_reduce(ratios, op=dist.ReduceOp.MAX, group=self.group)
                    eng.ratios(ratios, set=True)
                    eng.converge(n)
                for r0, r1 in slabs:
                    with torch.cuda.stream(comm):
                        gather(In[r0:r1])
                        arrived = torch.cuda.Event(enable_timing=time_comm)
                        arrived.record(comm)
                    compute.wait_event(arrived)
                    eng.source_rows(In, J, r0, r1)   # contraction of the next order on the slab that has landed
                if time_comm:
                    ev.append((swept, arrived))
            if (n - 1) % poll_every == 0:
                done = not any(r.active for r in eng.results())
        res = eng.results()
        allgather_columns(I, self.blocks, self.rank, self.group)
        if time_comm and ev:
            torch.cuda.synchronize()
            self.comm_ms = sum(a.elapsed_time(b) for a, b in ev)
        return I, res
