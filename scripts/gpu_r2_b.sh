#!/bin/bash
# Round 2, visit B: carry-free packed scan (row groups 2/4/8) against the two-op 16-bit kernel.
set -x
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2b_gpu_tests.log; tail -40 gpurun_out/r2b_gpu_tests.log
for g in 1 2 4 8; do
  timeout 200 python bench.py --workload cfg2 --qgroup $g --steps 5 --warmup 3 --parity-windows 2 > gpurun_out/r2b_cfg2_g$g.json 2> gpurun_out/r2b_cfg2_g$g.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2b_cfg2_g$g.json"))
    print("cfg2 G=$g ms/step %.2f scan %.2f tail %.2f bits %s slow %s two %s chk %.2e parity %s"%(j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["scan"]["operand_bits"],j["slow_path_samples"],j["scan"]["two_candidate_samples"],j["scan"]["check_max_rel_err"],j["parity_vs_reference_sample"]["pass"]))
except Exception as e:
    print("cfg2 G=$g failed", e); print(open("gpurun_out/r2b_cfg2_g$g.err").read()[-1500:])
PY
done
for g in 1 2 4 8; do
  timeout 300 python bench.py --workload cfg3 --samples 60000 --qgroup $g --steps 3 --warmup 1 --parity-windows 2 > gpurun_out/r2b_cfg3_g$g.json 2> gpurun_out/r2b_cfg3_g$g.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2b_cfg3_g$g.json"))
    print("cfg3/60000 G=$g ms/step %.2f scan %.2f tail %.2f bits %s slow %s two %s chk %.2e parity %s"%(j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["scan"]["operand_bits"],j["slow_path_samples"],j["scan"]["two_candidate_samples"],j["scan"]["check_max_rel_err"],j["parity_vs_reference_sample"]["pass"]))
except Exception as e:
    print("cfg3 G=$g failed", e); print(open("gpurun_out/r2b_cfg3_g$g.err").read()[-1500:])
PY
done
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 1 > gpurun_out/r2b_bench_cfg4_locate.json 2> gpurun_out/r2b_bench_cfg4_locate.err; tail -c 900 gpurun_out/r2b_bench_cfg4_locate.json; tail -5 gpurun_out/r2b_bench_cfg4_locate.err
