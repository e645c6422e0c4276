"""Device time of single solves of BASELINE configs 1-3 (default 800 x 1002 grid, HG(0.5) standing in for the Mie aerosol)
and of one bench batch, with and without the CUDA-graph order loop (SOS_B200_GRAPH=0).  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sos_b200 as sos  # noqa: E402
import bench  # noqa: E402

CASES = {
    "config1_eva_lambert": dict(surface="lambert", mu0=0.5, tauStar_atm=0.124, tauStar_aer=0.12, alb_aer=0.97, grd_alb=0.15),
    "config2_wildfire_lambert": dict(surface="lambert", mu0=0.5, tauStar_atm=0.124, tauStar_aer=0.0075, alb_aer=0.9, grd_alb=0.15, z_up=15.0, z_down=14.0),
    "config3_eva_specular": dict(surface="specular", mu0=0.5, tauStar_atm=0.124, tauStar_aer=0.12, alb_aer=0.97, grd_alb=0.15),
}


def timed(bs, reps=5):
    bs.solve(poll_every=2)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = bs.solve(poll_every=2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, int(r.n_orders.max())


out = {}
for graph in ("1", "0"):
    os.environ["SOS_B200_GRAPH"] = graph
    for name, kw in CASES.items():
        sc = sos.Scenario(nb_layers=800, nb_angles=501, atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.5), **kw)
        bs = sos.BatchSolver([sc])
        ms, n = timed(bs)
        out.setdefault(name, {})["graph" if graph == "1" else "no_graph"] = {"ms": ms, "orders": n, "generated_source": bool(bs.engine.generated_source)}
        bs.engine.close()
    bs = sos.BatchSolver(bench.make_scenarios(sos, 96))
    ms, n = timed(bs, 3)
    out.setdefault("bench_batch_96", {})["graph" if graph == "1" else "no_graph"] = {"ms": ms, "orders": n}
    bs.engine.close()
print(json.dumps(out))
