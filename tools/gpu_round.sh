#!/bin/bash
# One GPU-box pass (1 GPU): full GPU test suite, smoke, both bench arms, ncu launch list (full-size table) and `--set full` captures
# (4 Mi-row table: kernel replay saves / restores device memory — never run `--set full` against the 140 GB full-size process).
# Outputs under gpurun_out/; copy what is to be judged into profiles/.
cd "$(dirname "$0")/.."
O=gpurun_out
R=${ROUND:-r02}
timeout 900 python -m pytest tests -m gpu -q --tb=short > $O/${R}_pytest_gpu.log 2>&1; tail -4 $O/${R}_pytest_gpu.log
# the opt-in tests of the features added after the GPU budget of round 2 was spent (admission, cache tier, mixed dims): first hardware run
RECSYS_B200_UNVERIFIED_GPU_TESTS=1 timeout 600 python -m pytest tests -m gpu -q --tb=short -k "zz_" > $O/${R}_pytest_gpu_optin.log 2>&1; tail -4 $O/${R}_pytest_gpu_optin.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${R}_smoke.log 2>&1; tail -2 $O/${R}_smoke.log
timeout 900 python bench.py > $O/${R}_bench_line.json 2> $O/${R}_bench.err; tail -c 600 $O/${R}_bench_line.json; tail -3 $O/${R}_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/${R}_bench_ref.json 2> $O/${R}_bench_ref.err; tail -c 400 $O/${R}_bench_ref.json
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${R}_bench_launches.csv python bench.py --ncu --no-hstu --no-cpu --no-e2e > $O/ncu_launch.log 2>&1
  timeout 700 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"forward_seq|forward_pool|train_lookup_tile|train_evict|train_init_rows|unique_|backward_tiles|radix_scatter" -c 24 -f -o $O/${R}_demb_full python bench.py --ncu --capacity 4194304 --plaw-draws 16777216 --no-hstu --no-cpu --no-e2e > $O/ncu_full.log 2>&1
  ls -la $O/*.ncu-rep $O/${R}_*.csv
fi
