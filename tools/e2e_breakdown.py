"""Where the end-to-end step of bench.py goes: BatchSolver.update / solve / results with a device sync after each (resident plan)."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sos_b200 as sos  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
S = 96
batches = [bench.make_scenarios(sos, S, r) for r in range(4)]
bs = sos.BatchSolver(batches[0], device=dev)
for k in range(3):
    bs.update(batches[k % 4]); r = bs.solve(poll_every=2); bs.results(r, quadratures=True, fields=False)
torch.cuda.synchronize()
rows = []
for k in range(8):
    t0 = time.perf_counter(); bs.update(batches[k % 4]); t1 = time.perf_counter(); torch.cuda.synchronize(); t1s = time.perf_counter()
    r = bs.solve(poll_every=2); t2 = time.perf_counter(); torch.cuda.synchronize(); t2s = time.perf_counter()
    out = bs.results(r, quadratures=True, fields=False); t3 = time.perf_counter()
    rows.append((t1 - t0, t1s - t1, t2 - t1s, t2s - t2, t3 - t2s, t3 - t0))
a = np.array(rows) * 1e3
print("ms: update(host) %.2f | sync after update %.2f | solve(host returns) %.2f | sync after solve %.2f | results %.2f | total %.2f" % tuple(np.median(a, axis=0)))
# un-synced total (what bench measures)
ts = []
for k in range(8):
    t0 = time.perf_counter(); bs.update(batches[k % 4]); r = bs.solve(poll_every=2); out = bs.results(r, quadratures=True, fields=False); ts.append(time.perf_counter() - t0)
print("unsynced step ms: %.2f" % (np.median(ts) * 1e3))
pr = cProfile.Profile(); pr.enable()
for k in range(4):
    bs.update(batches[k % 4]); r = bs.solve(poll_every=2); out = bs.results(r, quadratures=True, fields=False)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
