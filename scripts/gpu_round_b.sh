#!/bin/bash
# Second GPU visit of round 1: the new stages (output stage, pre-processing), bench, launch lists.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 420 python -m pytest tests/test_gpu_output_stage.py tests/test_gpu_preprocess.py -q -m gpu 2>&1 | tail -60 > gpurun_out/new_tests.log; tail -25 gpurun_out/new_tests.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_output_stage.py -q -m gpu -x -k "8193 or channel_selection or quantised or decimate or envelope_matches_scipy_hilbert and 1001" 2>&1 | tail -15 > gpurun_out/sanitize.log; tail -8 gpurun_out/sanitize.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_b.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_sos_apply -s 2 -c 1 -o gpurun_out/prof_sos_apply python bench.py --steps 1 --warmup 1 --samples 12000 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_output_stage.py --deselect tests/test_gpu_preprocess.py 2>&1 | tail -8 > gpurun_out/rest_tests.log; tail -8 gpurun_out/rest_tests.log
ls -la gpurun_out | tail -15
