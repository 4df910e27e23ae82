#!/bin/bash
# Round 2, visit D: 16-bit offsets (MODE 11) vs 32-bit offsets (MODE 8) for the 15-bit packed scan.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_precision.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 > gpurun_out/r2d_gpu_tests.log; tail -8 gpurun_out/r2d_gpu_tests.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    j=json.load(open(f))
    print("%s ms/step %.2f scan %.2f tail %.2f bits %s slow %s two %s chk %.2e parity %s e2e %.4g"%(tag,j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["scan"]["operand_bits"],j["slow_path_samples"],j["scan"]["two_candidate_samples"],j["scan"]["check_max_rel_err"],(j.get("parity_vs_reference_sample") or {}).get("pass"),j["e2e"]["value"]))
except Exception as e:
    print(tag,"failed",e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
timeout 200 python bench.py --workload cfg2 --steps 5 --warmup 3 --parity-windows 2 > gpurun_out/r2d_cfg2_m11.json 2> gpurun_out/r2d_cfg2_m11.err; summ "cfg2 half-offsets" gpurun_out/r2d_cfg2_m11.json
QMB_WIDE_OFFSETS=1 timeout 200 python bench.py --workload cfg2 --steps 5 --warmup 3 --parity-windows 2 > gpurun_out/r2d_cfg2_m8.json 2> gpurun_out/r2d_cfg2_m8.err; summ "cfg2 wide-offsets" gpurun_out/r2d_cfg2_m8.json
timeout 600 python bench.py --steps 3 --warmup 1 > gpurun_out/r2d_cfg3_full.json 2> gpurun_out/r2d_cfg3_full.err; summ "cfg3 full" gpurun_out/r2d_cfg3_full.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/r2d_prof_scan_cfg2_m11 python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu_full.log 2>&1
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_ref_cfg3.json 2> gpurun_out/r2d_ref_cfg3.err; tail -c 500 gpurun_out/r2d_ref_cfg3.json
