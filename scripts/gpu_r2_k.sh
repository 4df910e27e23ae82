#!/bin/bash
# Round 2, visit K: warp-parallel look-back in the fused pre-processing pass.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_reference_adapter.py tests/test_iceland_locate.py tests/test_gpu_ranges.py tests/test_gpu_scan_api.py -q -m gpu 2>&1 > gpurun_out/r2k_gpu_tests.log; tail -8 gpurun_out/r2k_gpu_tests.log
for wl in cfg2 cfg3; do
  timeout 200 python bench.py --workload $wl --pre-only --steps 10 --warmup 3 > gpurun_out/r2k_pre_$wl.json 2> gpurun_out/r2k_pre_$wl.err; python -c "
import json; j=json.load(open('gpurun_out/r2k_pre_$wl.json')); print('$wl', j['kernel_ms'], j['roofline']['frac'])"; tail -3 gpurun_out/r2k_pre_$wl.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sos_fused -c 2 -o gpurun_out/r2k_prof_sos_fused python bench.py --workload cfg3 --pre-only --steps 1 --warmup 1 > gpurun_out/r2k_ncu_sos.log 2>&1
