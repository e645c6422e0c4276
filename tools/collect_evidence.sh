#!/bin/bash
# One-GPU evidence of a round (run through gpurun from the repo root): smoke, bench lines, single-solve times, the ncu launch
# list of one batch solve and the `ncu --set full` capture of the kernels of one scattering order.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/ev_smoke.log 2>&1; echo rc=$? >> $O/ev_smoke.log
python bench.py > $O/ev_bench_n1.json 2> $O/ev_bench_n1.err; echo rc=$? >> $O/ev_bench_n1.err
python bench.py --workload thick --steps 3 --warmup 3 > $O/ev_bench_thick_n1.json 2> $O/ev_bench_thick_n1.err
python tools/single_solve_times.py > $O/ev_single.json 2> $O/ev_single.err
# (ncu only after the same commands have exited 0 without it)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/ev_launches.csv python tools/order_bench.py 96 1 > $O/ev_ncu_launch.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sweep_local|sweep_carry|sweep_apply2|sweep_zone|jn_gemm_fold" -s 5 -c 5 -f -o $O/ev_order python tools/order_bench.py 96 1 > $O/ev_ncu_full.log 2>&1
echo done > $O/ev_done
