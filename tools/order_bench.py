"""Time the order loop of bench.py's batch (no e2e, no CPU leg): per-class CUDA-event times of one solve.
Usage: python tools/order_bench.py [S] [repeats]   (env SOS_B200_GENSRC=0 for the chunked scan)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sos_b200 as sos  # noqa: E402
import bench  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scen = bench.make_scenarios(sos, S)
bs = sos.BatchSolver(scen)
eng = bs.engine
lib = sos._lib.load()
print("gensrc", eng.gensrc_enabled, "generated", eng.generated_source, flush=True)
res = bs.solve(poll_every=2)
torch.cuda.synchronize()
norders = int(np.sum(res.n_orders - 1))
for _ in range(reps):
    lib.sos_set_profiling(eng._plan, 1)
    t0 = time.perf_counter()
    res = bs.solve(poll_every=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ms2 = (C.c_double * 4)()
    sp2 = (C.c_longlong * 4)()
    lib.sos_get_profile(eng._plan, ms2, sp2, None)
    lib.sos_set_profiling(eng._plan, 0)
    elems = norders * 800 * 1002
    print(f"solve {dt * 1e3:.2f} ms | contraction {ms2[0]:.2f} ms / {sp2[0]} launches (dense {ms2[3]:.2f}) | sweeps {ms2[1]:.2f} ms / {sp2[1]} launches (apply {ms2[2]:.2f}) "
          f"| sweeps: {32 * elems / (ms2[1] * 1e-3) * 1e-9:.0f} GB/s at 32 B/elem, {elems / (ms2[1] * 1e-3) * 1e-9:.1f} Gelem/s "
          f"| orders {int(res.n_orders.min())}..{int(res.n_orders.max())}", flush=True)
