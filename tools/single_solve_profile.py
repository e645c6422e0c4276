"""One default-grid single solve (BASELINE configs[2]: EVA aerosol, specular surface) a few times; run under
`ncu --metrics gpu__time_duration.sum --csv` for the per-kernel launch list of a single solve."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sos_b200 as sos  # noqa: E402

sc = sos.Scenario(surface="specular", mu0=0.5, tauStar_atm=0.124, tauStar_aer=0.12, alb_aer=0.97, grd_alb=0.15,
                  atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.5))
bs = sos.BatchSolver([sc])
for _ in range(3):
    r = bs.solve(poll_every=2)
torch.cuda.synchronize()
print("orders", int(r.n_orders[0]))
