for wt in 32 64; do for bm in 128 64; do
  echo "== warptile=$wt bm=$bm"
  SOS_GEMM_WARPTILE=$wt SOS_GEMM_BM=$bm timeout 60 python tools/bench_kernels.py --scenarios 96 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('S96', round(d['gemm_tflops_mean'],2))"
  SOS_GEMM_WARPTILE=$wt SOS_GEMM_BM=$bm timeout 60 python tools/bench_kernels.py --scenarios 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('S1', round(d['gemm_tflops_mean'],2), round(d['gemm_ms_mean']*1e3,1),'us')"
  SOS_GEMM_WARPTILE=$wt SOS_GEMM_BM=$bm timeout 60 python tools/bench_kernels.py --single --scenarios 1 --layers 10000 --angles 512 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4', round(d['gemm_tflops_mean'],2))"
done; done
