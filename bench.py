#!/usr/bin/env python
"""bench.py -- scattering-order updates/s of the SOS_AER hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scenarios S] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the critical-albedo style batch sweep): S independent
three-region scenarios per GPU on the reference's default grid (800 layers x 1002 mu), spanning
tau_aer x mu0 x omega_aer x surface albedo x aerosol phase function (HG / log-normal Mie stand-in / FWC); every GPU solves its own S
scenarios with no data-path collective (scaling: weak).  A "step" is one whole solve of the batch:
closed-form first order + the order loop to In/I < 1e-4 for every scenario.

  metric  = sum over scenarios and orders n >= 2 of L*N^2  /  time     
  value   : inputs already resident in HBM (plan, phase operands, coefficients uploaded before)
  e2e     : the public API call with HOST arrays in / NumPy out per step: plan creation, H2D of tau,
            coefficients and phase matrices, solve, D2H of the flux / diffusivity / heating-rate
            profiles, order counts and TOA net flux of every scenario (what a forcing sweep returns;
            the reference's SOS_Aer_radiative_forcing returns one float per solve)
  roofline: the dominant kernel class of the timed steps (CUDA events on the launching stream inside them): the four
            layer-sweep kernels (HBM bound, peak = MEASURED_PEAKS.json hbm_gbs) when the molecular rows of the
            contraction are low rank (Rayleigh: the default workload), with the contraction nested as
            roofline.contraction; otherwise the FP64 source contraction (jn_gemm_fold / jn_gemm_dmma) against the
            FP64 DMMA/DFMA throughput measured on this GPU in the same run (MEASURED_PEAKS.json has no FP64 entry),
            counting the FLOPs of the kernel that ran (the folded kernel needs half of the general one's)
  cpu_baseline / --impl reference: the NumPy oracle port of the reference algorithm
            (oracle/sos_oracle.py, method="slices") on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_DEFAULT, M_DEFAULT = 800, 501
WORKLOAD = "critical-albedo batch sweep (BASELINE configs[4]): S specular scenarios/GPU, 800x1002 grid"


def make_scenarios(sos, S, rank=0, L=L_DEFAULT, M=M_DEFAULT):
    """Deterministic sweep tau_aer x mu0 x omega_aer x albedo x phase."""
    taus = np.linspace(0.0075, 0.5, 10)
    mu0s = np.linspace(0.1, 1.0, 10)
    oms = np.linspace(0.7, 1.0, 10)
    albs = (0.05, 0.15, 0.3)
    # phase-function axis of BASELINE configs[4]: HG / Mie / FWC (Mie = the EVA log-normal mixture from the host stand-in)
    phases = (("hg", 0.5), ("mie_lognormal", sos.EVA_AEROSOL), ("fwc", 0.0))
    out = []
    for i in range(S):
        k = rank * S + i
        out.append(sos.Scenario(
            nb_layers=L, nb_angles=M, tauStar_atm=0.124,
            tauStar_aer=float(taus[k % 10]), mu0=float(mu0s[(k // 10 + 3 * k) % 10]),
            alb_aer=float(oms[(k // 100 + 7 * k) % 10]), grd_alb=float(albs[(k // 3 + k) % 3]),
            atm_phase=("rayleigh", 0.0), aer_phase=phases[k % 3], surface="specular"))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    nvidia-smi takes a few hundred ms to start, so the sampler is started before the warm-up and
    every line is stamped on arrival; stop(t0, t1) keeps the samples that fall inside [t0, t1]."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)  # let the last in-window sample arrive
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, r in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


_CPU_CACHE = {}


def cpu_port_inputs(L, M, seed_rank):
    """Inputs of one workload scenario for the NumPy port (built once per process, untimed)."""
    key = (L, M, seed_rank)
    if key not in _CPU_CACHE:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import sos_oracle as so
        import sos_b200 as sos
        scen = make_scenarios(sos, 1, seed_rank, L, M)[0]
        mu = so.mu_grid(M)
        P0a, Pa = sos.phase_matrices("rayleigh", M, mu, scen.mu0)
        P0e, Pe = sos.phase_matrices(scen.aer_phase[0], M, mu, scen.mu0, scen.aer_phase[1])
        sc = so.Scenario(mu0=scen.mu0, nb_layers=L, nb_angles=M, tauStar_atm=scen.tauStar_atm,
                         tauStar_aer=scen.tauStar_aer, grd_alb=scen.grd_alb, alb_aer=scen.alb_aer, surface="specular")
        tau, z, iu, idn = sc.geometry()
        I1 = so.first_order_regions(sc, tau, mu, iu, idn, P0a, P0e)
        lay = so.driver_layout(sc, tau, mu, iu, idn)
        _CPU_CACHE[key] = (so, sc, mu, iu, idn, Pa, Pe, lay, I1)
    return _CPU_CACHE[key]


def cpu_port_sample(orders=2, L=L_DEFAULT, M=M_DEFAULT, seed_rank=0):
    """One scenario of the workload through the NumPy port ("slices" = the reference's O(L^2 N)
    scheme): phase matrices and first order untimed, `orders` scattering orders timed.
    orders == 0 only builds the inputs.  Returns (units, seconds)."""
    so, sc, mu, iu, idn, Pa, Pe, lay, I1 = cpu_port_inputs(L, M, seed_rank)
    In = I1
    t0 = time.perf_counter()
    for _ in range(orders):
        J = so.source_regions(sc, In, mu, iu, idn, Pa, Pe)
        In = so.order_sweeps(lay, J, method="slices")
    dt = time.perf_counter() - t0
    return orders * L * (2 * M) ** 2, dt


def _cpu_worker(args):
    # every pool worker keeps ONE scenario of the workload (chosen by its worker index), so the inputs
    # built in the untimed first round are the ones the timed rounds reuse
    import multiprocessing as mp
    ident = mp.current_process()._identity
    orders, L, M, _ = args
    return cpu_port_sample(orders, L, M, ident[0] - 1 if ident else 0)


def run_reference(args):
    """--impl reference: the reference algorithm's CPU port on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    orders = 1
    L = args.cpu_layers
    with mp.get_context("spawn").Pool(procs) as pool:
        # untimed: imports, phase matrices and first order of every worker's scenario (chunksize 1 and as
        # many tasks as workers: each worker builds, and later reuses, its own scenario)
        pool.map(_cpu_worker, [(0, L, M_DEFAULT, i) for i in range(procs)], chunksize=1)
        times = []
        units = 0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(orders, L, M_DEFAULT, i) for i in range(procs)], chunksize=1)
            times.append(time.perf_counter() - t0)
            units = sum(r[0] for r in res)
    ms = 1e3 * float(np.mean(times))
    val = units / (ms * 1e-3)
    sample = (f"{procs} processes x 1 scenario x {orders} order(s) of the NumPy port (method='slices') on a "
              f"{L}x{2 * M_DEFAULT} grid per step; the port vectorises over mu what the reference loops over in Python")
    emit(({
        "impl": "reference", "metric": "scattering-order updates/s", "value": val, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "grid": [L, 2 * M_DEFAULT], "note": "CPU sample uses the workload's scenarios; "
                   "cost per order is O(L^2 N) in the reference scheme, so a reduced L flatters the CPU"},
        "cpu_baseline": {"value": val, "unit": "updates/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_thick(args, sos, torch, dist, dev, rank, world, W):
    """BASELINE configs[3]: optically thick FWC cloud layer (tau* = 30, omega = 0.9) on ONE large grid
    (10 000 layers x 1024 mu), run to In/I < 1e-4; for N > 1 the grid is sharded by mu blocks and the
    contraction reads the peers' I_n blocks by TMA over NVLink (strong scaling)."""
    L, M, tau_star, mu0, alb = 10000, 512, 30.0, 0.5, 0.9
    N = 2 * M
    mu = sos.mu_grid(M)
    tau = np.linspace(0, tau_star, L)
    w = sos.extrapolation_width(tau_star, M)
    coef = [sos.ScenarioCoefficients(mu0=mu0, grd_alb=0.0, tauStar_tot=tau_star, coef_atm=alb, extrap_width=(w, w, w))]
    eng = sos.SosEngine(mu, tau[None], coef, [0, L], sos._lib.SURFACE_NONE, device=dev)
    phase = "fwc"
    P, _ = eng.build_phase_matrix(phase)                      # built on the device
    eng.set_phase([P])
    Cc = np.zeros((1, 2, N))
    Cc[0, 0] = alb * sos.phase_P0(phase, M, mu, mu0)
    I1 = eng.first_order(Cc)
    blocks = sos.mu_blocks(N, M, world, M - w - 5)
    solver = sos.MuShardedSolver(eng, blocks, rank)
    peers = sos.PeerFields(eng, rank, world) if world > 1 else None

    def step():
        if world == 1:
            r = eng.solve(I1, max_orders=args.thick_orders, poll_every=8)
            return int(r.n_orders[0]), int(r.status[0])
        I, res = solver.solve_p2p(I1, peers, max_orders=args.thick_orders)
        return int(res[0].n_orders), int(res[0].status)

    for _ in range(W):
        n, status = step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    time.sleep(0.4)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    t0 = time.perf_counter()
    for k in range(args.steps):
        ev[2 * k].record()
        n, status = step()
        ev[2 * k + 1].record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    ms = float(np.mean([ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    units = (n - 1) * L * N * N
    if peers is not None:
        torch.cuda.synchronize(dev)
        dist.barrier()
        peers.close()
    if rank == 0:
        emit(({
            "metric": "scattering-order updates/s", "value": units / (ms * 1e-3), "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "thick FWC cloud layer (BASELINE configs[3]): one 10000x1024 grid, tau*=30, omega=0.9, "
                                   "to In/I<1e-4", "orders": n, "status": status, "ms_per_order": ms / max(n - 1, 1),
                       "sharding": "none" if world == 1 else "mu blocks of %d columns, contraction reads peer blocks by TMA over NVLink" % (N // world),
                       "contraction": "folded" if (eng.folded and world == 1) else "general",
                       "l2": "working set 4 x 82 MB fields + operand > 126 MB L2"},
            "clocks": clocks}))


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line, on the process's real stdout (see main: libraries may chat on fd 1)."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line: anything a library prints on fd 1 (NCCL prints its version there when
    # NCCL_DEBUG is set) goes to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scenarios", type=int, default=96, help="scenarios per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-layers", type=int, default=800)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "thick"],
                    help="sweep (default): BASELINE configs[4] batch; thick: configs[3], one 10000x1024 grid, mu-sharded for N>1")
    ap.add_argument("--thick-orders", type=int, default=300, help="order cap of the thick workload")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

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
    W = max(3, args.warmup)
    if args.workload == "thick":
        run_thick(args, sos, torch, dist, dev, rank, world, W)
        if world > 1:
            dist.destroy_process_group()
        return
    S, L, M = args.scenarios, L_DEFAULT, M_DEFAULT
    N = 2 * M

    scen = make_scenarios(sos, S, rank)
    bs = sos.BatchSolver(scen, device=dev)       # plan + phase operands resident
    # scenarios on which the reference itself would die with IndexError (blend-search overrun, Q11)
    # are not valid workload members: swap their aerosol phase function for HG(0.7) once, up front
    st = bs.solve(poll_every=2).status
    n_swapped = int(np.sum((st & 1) != 0))
    if n_swapped:
        import dataclasses
        scen = [dataclasses.replace(sc, aer_phase=("hg", 0.7)) if (st[i] & 1) else sc for i, sc in enumerate(scen)]
        bs.engine.close()
        bs = sos.BatchSolver(scen, device=dev)
    eng = bs.engine
    lib = sos._lib.load()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        res = bs.solve(poll_every=2)
        return res

    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(W):
        res = step_resident()
    units_per_step = int(np.sum(res.n_orders - 1)) * L * N * N
    n_orders = res.n_orders.copy()
    barrier()
    lib.sos_set_profiling(eng._plan, 1)
    l0 = eng.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    barrier()
    t_begin = time.perf_counter()
    t_wall = t_begin
    for k in range(args.steps):
        flush.zero_()                                  # L2 flush between timed iterations (untimed)
        ev[2 * k].record()
        step_resident()
        ev[2 * k + 1].record()
    barrier()
    t_end = time.perf_counter()
    t_wall = t_end - t_wall
    clocks = sampler.stop(t_begin, t_end)
    launches = eng.launches - l0
    import ctypes as C
    ms2 = (C.c_double * 4)()
    sp2 = (C.c_longlong * 4)()
    lib.sos_get_profile(eng._plan, ms2, sp2, None)
    lib.sos_set_profiling(eng._plan, 0)
    step_ms = float(np.mean([ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]))
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units_per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    step_ms_max, units_all = float(t.item()), float(u.item())
    value = units_all / (step_ms_max * 1e-3)

    # ---------------- roofline of the dominant kernel ----------------
    gemm_ms, gemm_launches = float(ms2[0]), int(sp2[0])
    sweep_ms, sweep_spans = float(ms2[1]), int(sp2[1])
    # FLOPs the kernel's algorithm needs: 2*L*N^2 per scenario-order for the general contraction;
    # the folded contraction (centrosymmetric operands, csrc/gemm_fold.cuh) computes the same J with two M x M
    # contractions per row = L*N^2 FLOP.  `achieved` counts what the shipped kernel's algorithm needs (no padding,
    # aerosol rows once), so frac stays a statement about the kernel; `value` keeps the SURVEY 8d unit.
    folded = bool(eng.folded)
    flops_general = 2.0 * units_per_step * args.steps
    flops = flops_general * (0.5 if folded else 1.0)
    achieved = flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
    pk = C.c_double()
    lib.sos_fp64_peak(1, 3, C.byref(pk))
    pk_dfma = C.c_double()
    lib.sos_fp64_peak(0, 3, C.byref(pk_dfma))
    peak = max(pk.value, pk_dfma.value)
    sweep_bytes = 32.0 * float(np.sum(n_orders - 1)) * L * N * args.steps
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (all 96 scenarios
    # active: one launch of order 2), scaled to the average number of active scenarios per timed launch
    traffic = None
    traffic_note = None
    tfile = "r01_ncu_fold_traffic.json" if folded else "r01_ncu_gemm_traffic.json"
    try:
        with open(os.path.join(ROOT, "profiles", tfile)) as f:
            tj = json.load(f)
        active_per_launch = float(np.sum(n_orders - 1)) * args.steps / max(gemm_launches, 1)
        traffic = tj["dram_bytes_per_active_scenario"] * active_per_launch
        traffic_note = ("dram__bytes_read+write per launch from profiles/%s (%.3e B at %d active "
                        "scenarios; algorithmic %.3e B) scaled to %.1f active scenarios per timed launch"
                        % (tfile, tj["dram_bytes_per_launch"], tj["scenarios"], tj["algorithmic_bytes_per_launch"], active_per_launch))
    except (OSError, KeyError, ValueError):
        pass
    roofline = {
        "bound": "tensor",
        "kernel": ("jn_gemm_fold_kernel (FP64 source contraction folded on the operand's centrosymmetry, DMMA m8n8k4)" if folded
                   else "jn_gemm_dmma_kernel (FP64 source contraction, DMMA m8n8k4)"),
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "traffic": traffic, "traffic_source": traffic_note,
        "algorithmic_flops_per_launch": flops / max(gemm_launches, 1),
        "flops_model": ("folded: L*N^2 FLOP per scenario-order (two M x M contractions per row); the general "
                        "contraction of SURVEY 8d needs 2*L*N^2" if folded else "2*L*N^2 FLOP per scenario-order (SURVEY 8d)"),
        "general_equivalent_tflops": flops_general / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else None,
        "peak_source": "FP64 DMMA m8n8k4 loop measured on this GPU in this run (sos_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 entry; DFMA loop measured %.1f TFLOP/s" % pk_dfma.value,
        "gemm_ms_per_launch": gemm_ms / max(gemm_launches, 1), "gemm_launches": gemm_launches,
        "gemm_share_of_step": gemm_ms / (step_ms * args.steps),
        "sweeps": {"bound": "hbm", "achieved": sweep_bytes / (sweep_ms * 1e-3) * 1e-9 if sweep_ms > 0 else None,
                   "unit": "GB/s", "algorithmic_bytes_per_element": 32, "ms_per_order": sweep_ms / max(sweep_spans, 1)},
    }
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hb = json.load(f).get("hbm_gbs")
        roofline["sweeps"]["peak"] = hb
        if hb and roofline["sweeps"]["achieved"]:
            roofline["sweeps"]["frac"] = roofline["sweeps"]["achieved"] / hb
    except OSError:
        roofline["sweeps"]["peak"] = 6650.0
        roofline["sweeps"]["peak_source"] = "fallback"
    lowrank = [int(r) for r in getattr(eng, "lowrank", [])]
    if any(lowrank):
        # Rows whose operand is low rank (Rayleigh: rank 2) no longer go through the DMMA kernel: the contraction is then a
        # mix of an HBM-bound skinny product (those rows) and the dense folded kernel (aerosol rows), and its FLOP count
        # against the tensor peak stops meaning anything -- report it by time and in the SURVEY 8d unit only.
        roofline["kernel"] = ("jn_lowrank_kernel (rows of low-rank operands, ranks %s: HBM bound) + jn_gemm_fold_kernel "
                              "(aerosol rows: FP64 DMMA)" % lowrank)
        roofline["flops_model"] = ("general_equivalent_tflops = 2*L*N^2 per scenario-order / time (SURVEY 8d); executed: 4*r*N FLOP per "
                                   "low-rank row + N^2 per dense folded row")
        roofline["achieved"] = roofline["frac"] = None
    if sweep_ms > gemm_ms:
        # the layer sweeps are now the dominant kernels of a step: they carry the headline roofline (HBM bound)
        sw = roofline.pop("sweeps")
        contraction = roofline
        straffic = None
        snote = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_ncu_sweeps_traffic.json")) as f:
                tj = json.load(f)
            active_per_order = float(np.sum(n_orders - 1)) * args.steps / max(sweep_spans, 1)
            straffic = tj["dram_bytes_per_active_scenario"] * active_per_order
            snote = ("dram__bytes_read+write of the four sweep kernels per order from profiles/r01_ncu_sweeps_traffic.json (%.3e B at "
                     "%d active scenarios; algorithmic %.3e B) scaled to %.1f active scenarios per timed order"
                     % (tj["dram_bytes_per_order"], tj["scenarios"], tj["algorithmic_bytes_per_order"], active_per_order))
        except (OSError, KeyError, ValueError):
            pass
        roofline = {
            "bound": "hbm",
            "kernel": "sweep_local + sweep_carry + sweep_apply + sweep_zone (layer sweeps, mu->0 rules, accumulate, convergence ratios)",
            "achieved": sw["achieved"], "peak": sw.get("peak"), "unit": "GB/s", "frac": sw.get("frac"),
            "traffic": straffic, "traffic_source": snote,
            "algorithmic_bytes_per_element": 32, "ms_per_order": sw["ms_per_order"],
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "peak_source" not in sw else sw["peak_source"],
            "sweeps_share_of_step": sweep_ms / (step_ms * args.steps),
            "contraction": contraction,
        }

    # ---------------- end to end through the public API (host arrays in, NumPy out) ----------------
    phases = sos.drivers._PHASES  # host-side phase matrices were built during plan creation above
    h2d = 0
    d2h = 0

    def step_e2e():
        nonlocal h2d, d2h
        b = sos.BatchSolver(scen, device=dev)          # plan creation + H2D of tau, coefficients, P
        r = b.solve(poll_every=2)
        out = b.results(r, quadratures=True, fields=False)   # D2H: flux/diffusivity/heating profiles, n, TOA net flux
        h2d = (b.tau.nbytes + b.Ccoef.nbytes + b.engine.h2d_phase_bytes + b.mu.nbytes)   # phase operands are uploaded once and stay resident
        d2h = sum(5 * o.flux_up.nbytes for o in out) + 40 * len(out)
        cnt = b.engine.launches
        b.engine.close()
        return cnt

    for _ in range(2):
        step_e2e()
    barrier()
    e2e_times = []
    e2e_launches = 0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_launches += step_e2e()
        torch.cuda.synchronize(dev)
        e2e_times.append(time.perf_counter() - t0)
    te = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = units_all / float(te.item())

    if rank == 0:
        line = {
            "metric": "scattering-order updates/s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": step_ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scenarios_per_gpu": S, "layers": L, "mu_columns": N,
                       "orders_per_scenario": [int(n_orders.min()), int(n_orders.max())],
                       "l2": "256 MB flush between timed steps; per-step working set %.0f MB > 126 MB L2" % (3 * S * L * eng.ld * 8 / 1e6),
                       "scenarios_swapped_for_blend_overrun": n_swapped,
                       "contraction": ("folded (centrosymmetric operands, defect %.1e)%s" % (eng.fold_defect, "; low-rank rows (ranks %s)" % [int(r) for r in eng.lowrank] if any(eng.lowrank) else "")) if eng.folded else "general",
                       "phase_functions": "HG(0.5) / log-normal Mie mixture (EVA aerosol, host Lorenz-Mie stand-in: miepython absent) / FWC table"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * float(te.item())},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu and world == 1:   # the CPU baseline leg is an N = 1 item
            units, dt = cpu_port_sample(orders=2)
            line["cpu_baseline"] = {
                "value": units / dt, "unit": "updates/s", "cores": 1, "kind": "port",
                "sample": "1 scenario of the workload, 2 scattering orders at the full 800x1002 grid through the NumPy "
                          "port of the reference algorithm (oracle method='slices', %.1f s); the unmodified reference "
                          "measured 3.0e7 updates/s on one core (BASELINE.md)" % dt}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
