#!/usr/bin/env python
"""Wall time of the single-scenario drivers (BASELINE configs 0-2) through the public API, host arrays out."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sos_b200 as sos

RUNS = {
    "config0 EVA Lambertian": (sos.SOS_Aer_main_lambertian, dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
    "config1 wildfire Lambertian": (sos.SOS_Aer_main_lambertian, dict(tauStar_atm=0.124, tauStar_aer=0.0075, z_up=15, z_down=14, grd_alb=0.15, alb_aer=0.97)),
    "config2 EVA specular": (sos.SOS_Aer_main_specular, dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
}
out = {}
for name, (fn, kw) in RUNS.items():
    kw = dict(kw, atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.5))
    for _ in range(3):
        r = fn(**kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); r = fn(**kw); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    # device-only order loop
    sc = sos.Scenario(surface="lambert" if fn is sos.SOS_Aer_main_lambertian else "specular", **kw)
    bs = sos.BatchSolver([sc])
    for _ in range(3):
        bs.solve()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        bs.solve()
    e1.record(); torch.cuda.synchronize()
    bs.engine.close()
    out[name] = {"orders": r.n, "api_ms_median": 1e3 * float(np.median(ts)), "device_solve_ms": e0.elapsed_time(e1) / 10,
                 "reference_minutes_measured_in_survey": "4-5 (27-32 s per order, SURVEY.md B.2)"}
print(json.dumps(out, indent=1))
