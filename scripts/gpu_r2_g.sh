#!/bin/bash
# Round 2, visit G: node-range slabs, fused pre-processing pass, gappy-trace path; full GPU suite.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 > gpurun_out/r2g_gpu_tests_full.log; tail -15 gpurun_out/r2g_gpu_tests_full.log
for leg in 0 1; do
  if [ $leg = 1 ]; then export QMB_PP_LEGACY=1; else unset QMB_PP_LEGACY; fi
  timeout 200 python bench.py --workload cfg2 --pre-only --steps 10 --warmup 3 > gpurun_out/r2g_pre_cfg2_legacy$leg.json 2> gpurun_out/r2g_pre_cfg2_legacy$leg.err; tail -c 700 gpurun_out/r2g_pre_cfg2_legacy$leg.json; tail -3 gpurun_out/r2g_pre_cfg2_legacy$leg.err
  timeout 200 python bench.py --workload cfg3 --pre-only --steps 10 --warmup 3 > gpurun_out/r2g_pre_cfg3_legacy$leg.json 2> gpurun_out/r2g_pre_cfg3_legacy$leg.err; tail -c 700 gpurun_out/r2g_pre_cfg3_legacy$leg.json; tail -3 gpurun_out/r2g_pre_cfg3_legacy$leg.err
done
unset QMB_PP_LEGACY
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r2g_cfg2.json 2> gpurun_out/r2g_cfg2.err; tail -c 400 gpurun_out/r2g_cfg2.json; tail -3 gpurun_out/r2g_cfg2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sos_fused -c 2 -o gpurun_out/r2g_prof_sos_fused python bench.py --workload cfg3 --pre-only --steps 1 --warmup 1 > gpurun_out/r2g_ncu_sos.log 2>&1
