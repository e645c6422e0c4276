import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import torch, numpy as np
import sos_b200 as sos, bench
dev = torch.device('cuda', 0)
scen = bench.make_scenarios(sos, 96, 0)
def step():
    b = sos.BatchSolver(scen, device=dev)
    r = b.solve(poll_every=2)
    out = b.results(r, quadratures=True, fields=False)
    b.engine.close()
    return out
step(); step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0=time.perf_counter(); step(); torch.cuda.synchronize(); t1=time.perf_counter()
pr.disable()
print('step s', t1-t0)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:5000])
import time as _t
for _ in range(3):
    t0=_t.perf_counter(); b = sos.BatchSolver(scen, device=dev); torch.cuda.synchronize(); t1=_t.perf_counter(); r = b.solve(poll_every=2); torch.cuda.synchronize(); t2=_t.perf_counter(); out = b.results(r, quadratures=True, fields=False); torch.cuda.synchronize(); t3=_t.perf_counter(); b.engine.close(); t4=_t.perf_counter()
    print(f'create {1e3*(t1-t0):.2f} solve {1e3*(t2-t1):.2f} results {1e3*(t3-t2):.2f} close {1e3*(t4-t3):.2f} ms')
