#!/usr/bin/env python
"""Experiment: one 96-scenario batch on one stream vs two 48-scenario half-batches on two streams
(driven by two host threads), to see whether the HBM-bound sweeps of one half overlap the
FP64-bound contraction of the other."""
import sys, time, threading
sys.path.insert(0, '.')
import torch, numpy as np
import sos_b200 as sos, bench

dev = torch.device('cuda', 0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
scen = bench.make_scenarios(sos, S, 0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.mean(ts))


one = sos.BatchSolver(scen, device=dev)
t_one = timed(lambda: one.solve(poll_every=2))
one.engine.close()

halves = [sos.BatchSolver(scen[i::2], device=dev) for i in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]


def run_half(i):
    with torch.cuda.stream(streams[i]):
        halves[i].solve(poll_every=2)


def both():
    th = [threading.Thread(target=run_half, args=(i,)) for i in range(2)]
    for t in th: t.start()
    for t in th: t.join()


t_two = timed(both)
t_seq = timed(lambda: (run_half(0), run_half(1)))
print(f"S={S}: one batch {t_one:.2f} ms | two halves sequential {t_seq:.2f} ms | two halves on two streams {t_two:.2f} ms")
