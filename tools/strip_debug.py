"""Small driver for the fused order kernel: one batch through strip.cuh and through the chunked scan, differences printed.
SOS_B200_SYNC_DEBUG=1 names the kernel that faults."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sos_b200 as sos  # noqa: E402

L = int(os.environ.get("DBG_L", 96))
M = int(os.environ.get("DBG_M", 251))
S = int(os.environ.get("DBG_S", 7))
surface = os.environ.get("DBG_SURF", "specular")
aer = (("hg", 0.5), ("fwc", 0.0), ("hg", 0.8))
scs = [sos.Scenario(nb_layers=L, nb_angles=M, mu0=(0.5, 0.23, 0.9, 1.0)[i % 4], tauStar_atm=(0.124, 0.05, 0.6, 0.3)[i % 4],
                    tauStar_aer=(0.12, 0.0, 0.9, 2.2)[i % 4], alb_aer=(0.97, 1.0, 0.8, 0.9)[i % 4], alb_atm=(1.0, 1.0, 0.9, 1.0)[i % 4],
                    grd_alb=(0.15, 0.0, 0.3, 0.8)[i % 4], atm_phase=("rayleigh", 0.0), aer_phase=aer[i % 3], surface=surface)
       for i in range(S)]
out = {}
os.environ["SOS_B200_STRIP_MIN"] = "1"
for flag in ("0", "1"):
    os.environ["SOS_B200_STRIP"] = flag
    bs = sos.BatchSolver(scs)
    print("strip", flag, "active", bs.engine.strip_active, "generated", bs.engine.generated_source, flush=True)
    res = bs.solve(keep_orders=3, max_orders=int(os.environ.get("DBG_ORDERS", 10000)))
    out[flag] = bs.results(res, quadratures=False, keep_orders=3)
    print("  n", [o.n for o in out[flag]], "status", res.status, flush=True)
    bs.engine.close()
for i, (a, b) in enumerate(zip(out["1"], out["0"])):
    den = np.max(np.abs(b.I))
    e = np.abs(a.I - b.I) / den
    t, m = np.unravel_index(np.argmax(e), e.shape)
    print(f"scen {i}: n {a.n}/{b.n}  max rel dI {e.max():.2e} at row {t} col {m}")
    for j in range(1, min(4, a.n, b.n)):
        ej = np.abs(a.I_saved[j] - b.I_saved[j]) / np.max(np.abs(b.I_saved[j]))
        t, m = np.unravel_index(np.argmax(ej), ej.shape)
        print(f"    order {j + 1}: {ej.max():.2e} at row {t} col {m}")
