"""cProfile of one single-scenario driver call through the public API (where do the host milliseconds go?)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sos_b200 as sos
kw = dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97, atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.5))
for _ in range(3):
    r = sos.SOS_Aer_main_specular(**kw)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); r = sos.SOS_Aer_main_specular(**kw); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("api ms", sorted(1e3 * t for t in ts))
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    r = sos.SOS_Aer_main_specular(**kw)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40); print(s.getvalue()[:7000])
