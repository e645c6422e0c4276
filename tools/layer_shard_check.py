#!/usr/bin/env python
"""Layer-block sharded solve of one large single-layer grid (BASELINE configs[3]) under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/layer_shard_check.py [--layers 10000] [--angles 512] [--orders 300] [--chunk-rows 0] [--check]

Rank 0 prints one JSON line: sharded time per solve / per order and (with --check) the deviation from the unsharded
solve of a plan with the same scan chunks on one GPU (expected: exactly 0) and the one-GPU time.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=10000)
    ap.add_argument("--angles", type=int, default=512)
    ap.add_argument("--tau", type=float, default=30.0)
    ap.add_argument("--orders", type=int, default=300, help="order cap")
    ap.add_argument("--phase", default="fwc")
    ap.add_argument("--chunk-rows", type=int, default=0, help="rows per scan chunk (0: the library's choice); the unsharded reference uses the same")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--emulate", default="", help="R/W: one process runs the kernels of rank R of W alone (set SOS_B200_PEER_TIMEOUT_MS=0); for ncu")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import sos_b200 as sos

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    shared_gpu = torch.cuda.device_count() < world   # fewer GPUs than ranks: the ranks share GPUs (time-sliced; correctness only)
    local %= torch.cuda.device_count()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if shared_gpu:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=dev)
    L, M = args.layers, args.angles
    N = 2 * M
    mu = sos.mu_grid(M)
    tau = np.linspace(0, args.tau, L)
    mu0, alb, g = 0.5, 0.9, 0.8
    P0 = sos.phase_P0(args.phase, M, mu, mu0, g)
    P = sos.phase_P(args.phase, M, mu, g)
    w = sos.extrapolation_width(args.tau, M)
    coef = [sos.ScenarioCoefficients(mu0=mu0, grd_alb=0.0, tauStar_tot=args.tau, coef_atm=alb, extrap_width=(w, w, w))]
    Cc = np.zeros((1, 2, N))
    Cc[0, 0] = alb * P0
    eng = sos.SosEngine(mu, tau[None], coef, [0, L], sos._lib.SURFACE_NONE, device=dev, chunk_rows=args.chunk_rows)
    eng.set_phase([P])
    I1 = eng.first_order(Cc)
    solver = sos.LayerShardedSolver(eng, rank, world, emulate=tuple(int(x) for x in args.emulate.split("/")) if args.emulate else None)

    for _ in range(2):
        solver.solve(I1, max_orders=args.orders)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(args.repeat):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        I, res = solver.solve(I1, max_orders=args.orders)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cpu" if shared_gpu else dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t[0]))
    ms = float(np.median(times))
    n_done = res[0].n_orders - 1
    out = {"world": world, "shared_gpu": bool(shared_gpu), "sharding": "layer blocks", "rows": solver.rows, "L": L, "N": N, "chunk_rows": args.chunk_rows, "folded": bool(eng.folded),
           "orders": int(n_done), "ms_total": ms, "ms_all": times, "ms_per_order": ms / max(n_done, 1), "status": int(res[0].status),
           "active": int(res[0].active), "updates_per_s": n_done * L * N * N / (ms * 1e-3)}
    if args.check:
        I_sh = I.view(-1, eng.ld)[:L, :N].cpu().numpy()
        if rank == 0:
            ref = sos.SosEngine(mu, tau[None], coef, [0, L], sos._lib.SURFACE_NONE, device=dev, chunk_rows=args.chunk_rows)
            ref.set_phase([P])
            J1 = ref.first_order(Cc)
            r = ref.solve(J1, max_orders=args.orders)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = ref.solve(J1, max_orders=args.orders)
            e1.record()
            torch.cuda.synchronize()
            out["one_gpu_ms"] = float(e0.elapsed_time(e1))
            out["speedup_vs_1gpu"] = out["one_gpu_ms"] / ms
            I_ref = r.I.view(-1, ref.ld)[:L, :N].cpu().numpy()
            out["max_rel_dev_vs_unsharded"] = float(np.max(np.abs(I_sh - I_ref)) / np.max(np.abs(I_ref)))
            out["bit_identical"] = bool(np.array_equal(I_sh, I_ref))
            out["orders_unsharded"] = int(r.n_orders[0]) - 1
            ref.close()
    if rank == 0:
        print(json.dumps(out))
    solver.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
