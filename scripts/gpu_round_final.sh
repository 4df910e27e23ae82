#!/bin/bash
# Round-1 closing visit: everything the driver will run, plus the profiles to commit.
set -x
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/gpu_tests.log; tail -6 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/prof_scan python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 200 python bench.py --workload cfg3 --samples 36000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 300 gpurun_out/bench_cfg3.json; tail -3 gpurun_out/bench_cfg3.err
timeout 100 python scripts/preprocess_bench.py > gpurun_out/preprocess_cfg3.json 2> gpurun_out/preprocess_cfg3.err; cat gpurun_out/preprocess_cfg3.json
ls -la gpurun_out | tail -14
