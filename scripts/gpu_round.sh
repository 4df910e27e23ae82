#!/bin/bash
# One GPU-box visit: full GPU test-suite, bench, ncu launch list + one full capture of k_scan.
set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/prof_scan python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
