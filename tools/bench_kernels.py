#!/usr/bin/env python
"""Time the hot kernels in isolation (all scenarios active) with CUDA events.

    python tools/bench_kernels.py [--scenarios S] [--layers L] [--angles M] [--reps R] [--single]

Used for roofline work and as the short command wrapped by `ncu --set full` (profiles/).
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
    ap.add_argument("--scenarios", type=int, default=96)
    ap.add_argument("--layers", type=int, default=800)
    ap.add_argument("--angles", type=int, default=501)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--single", action="store_true", help="one homogeneous layer (config 4 shape)")
    ap.add_argument("--tau", type=float, default=30.0)
    ap.add_argument("--chunk", type=int, default=0)
    args = ap.parse_args()
    import torch
    import sos_b200 as sos
    import bench

    dev = torch.device("cuda", 0)
    S, L, M = args.scenarios, args.layers, args.angles
    N = 2 * M
    if args.single:
        mu = sos.mu_grid(M)
        tau = np.tile(np.linspace(0, args.tau, L), (S, 1))
        w = sos.extrapolation_width(args.tau, M)
        coef = [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.0, tauStar_tot=args.tau, coef_atm=0.9, extrap_width=(w, w, w))
                for _ in range(S)]
        eng = sos.SosEngine(mu, tau, coef, [0, L], sos._lib.SURFACE_NONE, device=dev, chunk_rows=args.chunk)
        P0 = sos.phase_P0("hg", M, mu, 0.5, 0.8)
        eng.set_phase([sos.phase_P("hg", M, mu, 0.8)])
        Cc = np.zeros((S, 2, N))
        Cc[:, 0] = 0.9 * P0
        I1 = eng.first_order(Cc)
    else:
        scen = bench.make_scenarios(sos, S, 0, L, M)
        bs = sos.BatchSolver(scen, device=dev, chunk_rows=args.chunk)
        eng = bs.engine
        I1 = bs.first_order()
    eng.reset(I1)
    J = eng.new_field(zero=True)
    In = eng.new_field(zero=True)
    I = I1.clone()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts)), float(np.min(ts))

    g_mean, g_min = timeit(lambda: eng.source(I1, out=J), args.reps)
    s_mean, s_min = timeit(lambda: eng.sweeps(J, out=In, accumulate_into=I), args.reps)
    flops = 2.0 * S * L * N * N
    out = {
        "S": S, "L": L, "N": N, "gemm_kind": "folded" if eng.folded else "general",
        "gemm_executed_tflops_mean": flops * (0.5 if eng.folded else 1.0) / g_mean * 1e-9,
        "gemm_ms_mean": g_mean, "gemm_ms_min": g_min, "gemm_tflops_mean": flops / g_mean * 1e-9,
        "gemm_tflops_best": flops / g_min * 1e-9,
        "sweeps_ms_mean": s_mean, "sweeps_ms_min": s_min,
        "sweeps_algorithmic_gbs": 32.0 * S * L * N / s_mean * 1e-6,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
