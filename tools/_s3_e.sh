set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/ev_pytest_gpu.log 2>&1; echo rc=$? >> $O/ev_pytest_gpu.log
bash tools/collect_evidence.sh
python bench.py --impl reference --steps 2 --warmup 1 > $O/ev_bench_reference_arm.json 2> $O/ev_bench_reference_arm.err
echo done
