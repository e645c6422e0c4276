"""world_size-2 gloo tests of the multi-GPU host logic (no GPU): partitioning helpers, the column
all-gather used by mu-block sharding, the row gather used by layer-block sharding, the result gather used by scenario
sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sos_b200 as sos


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, M, rows):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- mu-block all-gather: every rank owns one column block of a (rows, ld) field ----
        blocks = sos.mu_blocks(N, M, world, zone_lo=M - 40)
        ld = (N + 15) // 16 * 16
        full = torch.arange(rows * ld, dtype=torch.float64).reshape(rows, ld)
        mine = torch.full((rows, ld), -1.0, dtype=torch.float64)
        c0, c1 = blocks[rank]
        mine[:, c0:c1] = full[:, c0:c1]
        sos.allgather_columns(mine, blocks, rank)
        assert torch.equal(mine[:, :N], full[:, :N]), "column all-gather did not rebuild the field"
        # ---- layer-block sharding: the final gather of the (uneven) row blocks of I ----
        class _Eng:   # what LayerShardedSolver.gather_rows needs of an engine
            L, ld = 3 * rows + 5, (N + 15) // 16 * 16
        cut = _Eng.L // 2 - 3
        solver = object.__new__(sos.LayerShardedSolver)
        solver.eng, solver.rank, solver.world, solver.group = _Eng, rank, world, None
        solver.rows = [(0, cut), (cut, _Eng.L)]
        solver._stage = None
        whole = torch.arange(_Eng.L * _Eng.ld, dtype=torch.float64).reshape(_Eng.L, _Eng.ld)
        part = torch.full_like(whole, -7.0)
        a, b = solver.rows[rank]
        part[a:b] = whole[a:b]
        solver.gather_rows(part)
        assert torch.equal(part, whole), "row gather did not rebuild the field"
        # ---- convergence ratios: MAX all-reduce ----
        r = torch.tensor([[0.1 * (rank + 1), 0.5 - 0.1 * rank]], dtype=torch.float64)
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
        assert torch.allclose(r, torch.tensor([[0.1 * world, 0.5]], dtype=torch.float64))
        # ---- scenario sharding: deal, "solve", gather in the original order ----
        scen = [sos.Scenario(mu0=0.1 + 0.05 * i) for i in range(7)]
        local = sos.shard_scenarios(scen, rank, world)
        results = [("solved", sc.mu0, rank) for sc in local]
        out = sos.gather_scenario_results(results, len(scen), rank, world)
        if rank == 0:
            assert [o[1] for o in out] == [sc.mu0 for sc in scen]
            assert [o[2] for o in out] == [i % world for i in range(len(scen))]
        else:
            assert out is None
        # ---- bench.py's second deal of a multi-GPU job: every rank reports the orders to convergence of its members of the
        #      first deal, every rank computes the same second deal (same make-up of order counts on every rank) ----
        import bench
        S = 12
        job, deals = bench.job_deal(sos, S, world, 40, 21)
        truth = np.array([7 + (5 * i) % 11 for i in range(world * S)])          # "measured" orders of the job's members
        mine, ov = bench.redeal_by_orders(sos, S, world, rank, truth[deals[rank]], np.zeros(S, dtype=np.int64), dist, torch,
                                          torch.device("cpu"), 40, 21)
        _, want = bench.job_deal(sos, S, world, 40, 21, cost=truth)
        key = lambda sc: (sc.tauStar_aer, sc.mu0, sc.alb_aer, sc.grd_alb, sc.aer_phase[0])
        assert [key(sc) for sc in mine] == [key(job[i]) for i in want[rank]] and not ov.any()
        tot = torch.tensor([int(truth[want[rank]].sum())])
        both = [torch.empty_like(tot) for _ in range(world)]
        dist.all_gather(both, tot)
        assert abs(int(both[0]) - int(both[1])) <= 4, "the second deal left the ranks with different work"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,M", [(1024, 512), (2048, 1024)])
def test_two_rank_collectives_gloo(N, M):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, N, M, 37), nprocs=2, join=True)


def test_partition_helpers():
    for n, w in ((10, 3), (7, 8), (96, 8), (1000, 7)):
        spans = [sos.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    scen = list(range(11))
    parts = [sos.shard_scenarios(scen, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == scen
    # config 4: 1024 columns over 8 GPUs -> 128-column blocks; zones stay whole
    b = sos.mu_blocks(1024, 512, 8, zone_lo=512 - 30 - 5)
    assert b[3] == (384, 512) and b[4] == (512, 640) and all((c1 - c0) % 128 == 0 for c0, c1 in b)
    with pytest.raises(ValueError):
        sos.mu_blocks(1024, 448, 8, zone_lo=400)      # a cut at 384..448 would split the downward zone
    with pytest.raises(ValueError):
        sos.mu_blocks(256, 128, 4, zone_lo=100)       # 64-column blocks: not a multiple of the GEMM tile
