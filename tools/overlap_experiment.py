#!/usr/bin/env python
"""Do the FP64-bound contraction (stream A) and the HBM-bound sweeps (stream B) overlap on one GPU?"""
import sys, time
sys.path.insert(0, '.')
import torch, numpy as np
import sos_b200 as sos, bench
dev = torch.device('cuda', 0)
S = 48
A = sos.BatchSolver(bench.make_scenarios(sos, S, 0), device=dev)
B = sos.BatchSolver(bench.make_scenarios(sos, S, 1), device=dev)
IA, IB = A.first_order().clone(), B.first_order().clone()
A.engine.reset(IA); B.engine.reset(IB)
JA, JB = A.engine.new_field(True), B.engine.new_field(True)
InB = B.engine.new_field(True)
B.engine.source(IB, out=JB)
sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
R = 10
def gemms():
    with torch.cuda.stream(sa):
        for _ in range(R): A.engine.source(IA, out=JA)
def sweeps():
    with torch.cuda.stream(sb):
        for _ in range(R): B.engine.sweeps(JB, out=InB)
def t(fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0)
tg, ts = t(gemms), t(sweeps)
tb = t(lambda: (gemms(), sweeps()))
print(f"{R} contractions {tg:.2f} ms | {R} sweeps {ts:.2f} ms | both on two streams {tb:.2f} ms (sum {tg+ts:.2f})")

# control: a plain torch elementwise kernel (HBM-bound) on stream B against the contractions on stream A
x = torch.zeros(128 * 1024 * 1024, dtype=torch.float64, device=dev)
def elem():
    with torch.cuda.stream(sb):
        for _ in range(R): x.mul_(1.0000001)
te = t(elem)
tbe = t(lambda: (gemms(), elem()))
print(f"{R} contractions {tg:.2f} ms | {R} torch mul_ {te:.2f} ms | both {tbe:.2f} ms (sum {tg+te:.2f})")
# control 2: only sweep_local-like traffic: sweeps vs elementwise (two HBM-bound streams)
tse = t(lambda: (sweeps(), elem()))
print(f"sweeps + mul_ on two streams {tse:.2f} ms (sum {ts+te:.2f})")

# control 3: single-CTA spin kernel on stream B (needs almost no resources)
def spin():
    with torch.cuda.stream(sb):
        for _ in range(R): torch.cuda._sleep(600000)
tsp = t(spin)
tbs = t(lambda: (gemms(), spin()))
print(f"contractions + _sleep: {tbs:.2f} ms (contractions {tg:.2f}, sleep {tsp:.2f})")
# control 4: cuBLAS fp64 matmul on stream A vs mul_ on stream B (torch-only concurrency in this harness)
a64 = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
def mm():
    with torch.cuda.stream(sa):
        for _ in range(R): torch.mm(a64, a64)
tm = t(mm)
tme = t(lambda: (mm(), elem()))
print(f"torch.mm fp64 {tm:.2f} ms | mul_ {te:.2f} ms | both {tme:.2f} ms (sum {tm+te:.2f})")
tms = t(lambda: (mm(), spin()))
print(f"torch.mm fp64 + _sleep: {tms:.2f} ms")
