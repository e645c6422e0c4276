import sys, os, time
sys.path.insert(0, '/root/repo')
import torch, sos_b200 as sos, bench
bs = sos.BatchSolver(bench.make_scenarios(sos, 96))
for _ in range(3): bs.first_order()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): bs.first_order()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("SOS_B200_FO_ROWS"), e0.elapsed_time(e1) / 10, "ms per first order")
