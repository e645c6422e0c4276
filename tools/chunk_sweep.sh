for ch in 0 24 34 48 64 100; do
  a=$(timeout 60 python tools/bench_kernels.py --scenarios 1 --chunk $ch 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['sweeps_ms_mean']*1e3,1))")
  b=$(timeout 60 python tools/bench_kernels.py --scenarios 16 --chunk $ch 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['sweeps_ms_mean']*1e3,1))")
  echo "chunk=$ch  S1 sweeps ${a}us  S16 sweeps ${b}us"
done
for ch in 0 128 256 512; do
  c=$(timeout 60 python tools/bench_kernels.py --single --scenarios 1 --layers 10000 --angles 512 --chunk $ch 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['sweeps_ms_mean']*1e3,1))")
  echo "cfg4 chunk=$ch sweeps ${c}us"
done
