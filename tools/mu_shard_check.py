#!/usr/bin/env python
"""mu-block sharded solve of one large single-layer grid (BASELINE config 4) under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/mu_shard_check.py [--layers 10000] [--angles 512] [--orders 24] [--check]

Rank 0 prints one JSON line: sharded time per order, all-gather share, and (with --check, G <= small
grids) the deviation from the unsharded solve on one GPU.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=10000)
    ap.add_argument("--angles", type=int, default=512)
    ap.add_argument("--tau", type=float, default=30.0)
    ap.add_argument("--orders", type=int, default=24)
    ap.add_argument("--phase", default="hg")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--slabs", type=int, default=1)
    ap.add_argument("--p2p", action="store_true", help="fused all-gather: contraction reads peer memory over NVLink")
    ap.add_argument("--ref-general", action="store_true", help="--check against the unsharded solve with the GENERAL contraction kernel "
                    "(what the sharded plans run): the two must then agree bit for bit; the default reference uses the folded kernel")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import sos_b200 as sos

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L, M = args.layers, args.angles
    N = 2 * M
    mu = sos.mu_grid(M)
    tau = np.linspace(0, args.tau, L)
    mu0, alb = 0.5, 0.9
    g = 0.8
    P0 = sos.phase_P0(args.phase, M, mu, mu0, g)
    P = sos.phase_P(args.phase, M, mu, g)
    w = sos.extrapolation_width(args.tau, M)
    coef = [sos.ScenarioCoefficients(mu0=mu0, grd_alb=0.0, tauStar_tot=args.tau, coef_atm=alb, extrap_width=(w, w, w))]
    Cc = np.zeros((1, 2, N))
    Cc[0, 0] = alb * P0
    eng = sos.SosEngine(mu, tau[None], coef, [0, L], sos._lib.SURFACE_NONE, device=dev)
    eng.set_phase([P])
    I1 = eng.first_order(Cc)
    blocks = sos.mu_blocks(N, M, world, M - w - 5)
    solver = sos.MuShardedSolver(eng, blocks, rank)
    peers = sos.PeerFields(eng, rank, world) if (args.p2p and world > 1) else None

    def run(max_orders, timed=False):
        if peers is not None:
            return solver.solve_p2p(I1, peers, max_orders=max_orders)
        return solver.solve(I1, max_orders=max_orders, time_comm=timed, row_chunks=args.slabs)

    for _ in range(2):
        run(8)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    I, res = run(args.orders + 1, timed=True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1), solver.comm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    n_done = res[0].n_orders - 1
    out = {"world": world, "p2p": bool(peers is not None), "slabs": args.slabs, "L": L, "N": N, "orders": int(n_done), "ms_total": float(ms[0]), "ms_per_order": float(ms[0]) / max(n_done, 1),
           "allgather_ms_per_order": float(ms[1]) / max(n_done, 1), "status": int(res[0].status),
           "updates_per_s": n_done * L * N * N / (float(ms[0]) * 1e-3)}
    if args.check:
        I_sh = I[:, :N].cpu().numpy()
        if rank == 0:
            ref = sos.SosEngine(mu, tau[None], coef, [0, L], sos._lib.SURFACE_NONE, device=dev, fold=False if args.ref_general else None)
            ref.set_phase([P])
            r = ref.solve(ref.first_order(Cc), max_orders=args.orders + 1)
            out["reference_contraction"] = "folded" if ref.folded else "general"
            I_ref = r.I[:, :N].cpu().numpy()
            out["max_rel_dev_vs_unsharded"] = float(np.max(np.abs(I_sh - I_ref)) / np.max(np.abs(I_ref)))
            out["orders_unsharded"] = int(r.n_orders[0]) - 1
    if rank == 0:
        print(json.dumps(out))
    if peers is not None:
        torch.cuda.synchronize()
        dist.barrier()
        peers.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
