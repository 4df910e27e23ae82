#!/bin/bash
# Round 2, visit C: default 15-bit packed scan; full GPU tests; ncu launch list + full capture of k_scan.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 > gpurun_out/r2c_gpu_tests_full.log; tail -15 gpurun_out/r2c_gpu_tests_full.log
for g in 2 4; do
  timeout 200 python bench.py --workload cfg2 --qgroup $g --steps 5 --warmup 3 --parity-windows 2 > gpurun_out/r2c_cfg2_g$g.json 2> gpurun_out/r2c_cfg2_g$g.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2c_cfg2_g$g.json"))
    print("cfg2 G=$g ms/step %.2f scan %.2f tail %.2f bits %s slow %s two %s chk %.2e parity %s"%(j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["scan"]["operand_bits"],j["slow_path_samples"],j["scan"]["two_candidate_samples"],j["scan"]["check_max_rel_err"],j["parity_vs_reference_sample"]["pass"]))
except Exception as e:
    print("cfg2 G=$g failed", e); print(open("gpurun_out/r2c_cfg2_g$g.err").read()[-1500:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/r2c_prof_scan_cfg2_g2 python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/r2c_prof_scan_cfg2_g4 python bench.py --workload cfg2 --qgroup 4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_full_g4.log 2>&1
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 1 > gpurun_out/r2c_bench_cfg4_locate.json 2> gpurun_out/r2c_bench_cfg4_locate.err; python -c "
import json; j=json.load(open('gpurun_out/r2c_bench_cfg4_locate.json')); print(j['ms_per_step'], j['stages'], j['e2e'])"
ls -la gpurun_out | tail -12
