import sys, time
sys.path.insert(0, '.')
import torch, numpy as np
import sos_b200 as sos, bench
dev = torch.device('cuda', 0)
scen = bench.make_scenarios(sos, 96, 0)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = sos.BatchSolver(scen, device=dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    b.engine.close()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  close {1e3*(t2-t1):.2f} ms")
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); b = sos.BatchSolver(scen, device=dev); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(12); print(s.getvalue()[:3000])
