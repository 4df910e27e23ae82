#!/bin/bash
# Round 2, visit A: new scan driver + device-resident locate + cfg3 default bench.
set -x
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2a_gpu_tests.log; tail -25 gpurun_out/r2a_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 1 > gpurun_out/r2a_bench_cfg3_n1.json 2> gpurun_out/r2a_bench_cfg3_n1.err; tail -c 1500 gpurun_out/r2a_bench_cfg3_n1.json; tail -5 gpurun_out/r2a_bench_cfg3_n1.err
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 1 > gpurun_out/r2a_bench_cfg4_locate.json 2> gpurun_out/r2a_bench_cfg4_locate.err; tail -c 1500 gpurun_out/r2a_bench_cfg4_locate.json; tail -5 gpurun_out/r2a_bench_cfg4_locate.err
timeout 200 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg2_n1.json 2> gpurun_out/r2a_bench_cfg2_n1.err; tail -c 600 gpurun_out/r2a_bench_cfg2_n1.json; tail -5 gpurun_out/r2a_bench_cfg2_n1.err
ls -la gpurun_out | tail -8
