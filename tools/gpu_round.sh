#!/bin/bash
# One GPU-box pass: full GPU test suite, smoke, both bench arms, ncu launch list + full captures (outputs under gpurun_out/).
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench_r01.json 2> $O/bench_r01.err; tail -c 1500 $O/bench_r01.json; tail -3 $O/bench_r01.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref_r01.json 2> $O/bench_ref_r01.err; tail -c 600 $O/bench_ref_r01.json
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r01_launches_bench.csv python bench.py --ncu --no-hstu --no-cpu > $O/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"backward_tiles|backward_windows|forward_seq|train_lookup|train_insert|unique_" -c 24 -f -o $O/r01_demb_full python bench.py --ncu --no-hstu --no-cpu > $O/ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:hstu -f -o $O/r01_hstu_full python tools/prof_hstu.py 8 > $O/ncu_hstu.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r01_launches_hstu.csv python tools/prof_hstu.py 32 > $O/ncu_hstu_launch.log 2>&1
  ls -la $O/*.ncu-rep $O/r01_*.csv
fi
