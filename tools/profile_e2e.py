import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import torch, numpy as np
import sos_b200 as sos, bench
dev = torch.device('cuda', 0)
scen = bench.make_scenarios(sos, 96, 0)
def step():
    b = sos.BatchSolver(scen, device=dev)
    r = b.solve(poll_every=2)
    out = b.results(r, quadratures=True)
    b.engine.close()
    return out
step(); step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0=time.perf_counter(); step(); torch.cuda.synchronize(); t1=time.perf_counter()
pr.disable()
print('step s', t1-t0)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
