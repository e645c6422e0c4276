#!/usr/bin/env python
"""Print the actual GPU-vs-reference deviations on the default-grid fixtures (tests only assert < 1e-10)."""
import ast, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sos_b200 as sos

def relmax(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))

RUNS = {
    "eva_spec": ("specular", dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
    "eva_lamb": ("lambertian", dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
    "wildfire_lamb": ("lambertian", dict(tauStar_atm=0.124, tauStar_aer=0.0075, z_up=15, z_down=14, grd_alb=0.15, alb_aer=0.97)),
    "shipped_spec": ("specular", {}),
}
print("contraction:", "general kernel (SOS_B200_FOLD=0)" if os.environ.get("SOS_B200_FOLD") == "0" else "folded kernel (default)")
for tag, (kind, kw) in RUNS.items():
    d = np.load(os.path.join(ROOT, "tests", "golden", f"default_{tag}.npz"))
    fn = sos.SOS_Aer_main_lambertian if kind == "lambertian" else sos.SOS_Aer_main_specular
    n_ref = int(d["n"])
    r = fn(keep_orders=n_ref, atm_phase=("rayleigh", 0.5), aer_phase=("hg", 0.5), **kw)
    rows = d["rows"]
    per_order = max(relmax(r.I_saved[j][rows], d["order_rows"][j]) for j in range(n_ref))
    print(f"{tag:14s} n={r.n} (ref {n_ref})  I {relmax(r.I[::40], d['I_sub']):.2e}  per-order max {per_order:.2e}  "
          f"flux_up {relmax(r.flux_up, d['flux_up']):.2e}  flux_down {relmax(r.flux_down, d['flux_down']):.2e}  "
          f"heating {np.max(np.abs(r.heating_rate - d['heating_rate'])) / np.max(np.abs(d['heating_rate'])):.2e}")
d = np.load(os.path.join(ROOT, "tests", "golden", "thick_fwc.npz"))
L, M, ts, mu0, alb = d["params"]; L, M = int(L), int(M)
mu = sos.mu_grid(M); tau = np.linspace(0, ts, L)
P0, P = sos.phase_matrices("fwc", M, mu, mu0)
w = sos.extrapolation_width(ts, M)
eng = sos.SosEngine(mu, tau[None], [sos.ScenarioCoefficients(mu0=mu0, grd_alb=0.0, tauStar_tot=ts, coef_atm=alb, extrap_width=(w, w, w))], [0, L], 0)
eng.set_phase([P]); Cc = np.zeros((1, 2, 2 * M)); Cc[0, 0] = alb * P0
res = eng.solve(eng.first_order(Cc))
print(f"thick_fwc      n={int(res.n_orders[0])} (ref {int(d['n'])})  I {relmax(eng.to_host(res.I)[::25], d['I_sub']):.2e}")
