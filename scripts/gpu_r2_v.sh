#!/bin/bash
# Round 2, visit V (1 GPU): full GPU suite with the recycled device buffers + the final 1-GPU bench lines.
set -x
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1800 python -m pytest tests -q -m gpu 2>&1 > gpurun_out/r2v_gpu_tests_full.log; tail -6 gpurun_out/r2v_gpu_tests_full.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2v_bench_n1_cfg3.json 2> gpurun_out/r2v_bench_n1_cfg3.err; tail -c 300 gpurun_out/r2v_bench_n1_cfg3.json; tail -3 gpurun_out/r2v_bench_n1_cfg3.err
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/r2v_bench_n1_cfg2.json 2> gpurun_out/r2v_bench_n1_cfg2.err; tail -c 300 gpurun_out/r2v_bench_n1_cfg2.json; tail -3 gpurun_out/r2v_bench_n1_cfg2.err
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 2 > gpurun_out/r2v_bench_n1_cfg4_locate.json 2> gpurun_out/r2v_bench_n1_cfg4_locate.err; tail -c 300 gpurun_out/r2v_bench_n1_cfg4_locate.json; tail -3 gpurun_out/r2v_bench_n1_cfg4_locate.err
ls -la gpurun_out | grep r2v
