"""Experiment: the 96-scenario bench batch as ONE plan against TWO half batches solved concurrently on two streams by two host
threads (ctypes releases the GIL): do the latency-bound tails of the order loops (few scenarios still iterating) overlap?
Usage: python tools/two_batch_experiment.py [S] [split: interleave|sorted]"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sos_b200 as sos  # noqa: E402
import bench  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
mode = sys.argv[2] if len(sys.argv) > 2 else "interleave"
scen = bench.make_scenarios(sos, S)
dev = torch.device("cuda", 0)


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), ts


one = sos.BatchSolver(scen, device=dev)
ms1, all1 = timed(lambda: one.solve(poll_every=2))
res = one.solve(poll_every=2)
orders = np.asarray(res.n_orders)
print("one plan: %.2f ms" % ms1, ["%.2f" % t for t in all1], "orders", int(orders.min()), int(orders.max()), flush=True)
one.engine.close()

if mode == "sorted":      # long-running scenarios in one half, short ones in the other
    idx = np.argsort(-orders, kind="stable")
    halves = [list(idx[: S // 2]), list(idx[S // 2:])]
else:
    halves = [list(range(0, S, 2)), list(range(1, S, 2))]
streams = [torch.cuda.Stream(device=dev) for _ in halves]
solvers = []
for h, st in zip(halves, streams):
    with torch.cuda.stream(st):
        solvers.append(sos.BatchSolver([scen[i] for i in h], device=dev))
torch.cuda.synchronize()


def run_half(k):
    with torch.cuda.stream(streams[k]):
        solvers[k].solve(poll_every=2)


def both():
    th = [threading.Thread(target=run_half, args=(k,)) for k in range(len(halves))]
    for t in th:
        t.start()
    for t in th:
        t.join()


ms2, all2 = timed(both)
print("two half plans on two streams (%s): %.2f ms" % (mode, ms2), ["%.2f" % t for t in all2], flush=True)
for k in range(len(halves)):
    msk, _ = timed(lambda: run_half(k))
    print("  half %d alone: %.2f ms" % (k, msk), flush=True)
