#!/bin/bash
# Round 2, visit M (1 GPU): slow-path subset filter -- parity tests, cfg3 / cfg5slab tails with and without it,
# DRAM traffic of cfg3's scan kernel (metrics-only ncu pass: a --set full capture of a 6 s kernel replays for minutes).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_precision.py tests/test_gpu_sizes.py tests/test_gpu_multirank.py tests/test_gpu_ranges.py -q -m gpu -x 2>&1 > gpurun_out/r2m_gpu_tests.log; tail -8 gpurun_out/r2m_gpu_tests.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    line=[l for l in open(f).read().splitlines() if l.startswith("{")][-1]
    j=json.loads(line)
    print("%s value %.4g ms/step %.2f scan %.2f tail %.2f slow %s e2e %.4g parity %s"%(tag,j["value"],j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["slow_path_samples"],j["e2e"]["value"],j.get("parity_vs_reference_sample",{}).get("pass")))
    print("   scan", j.get("scan"))
except Exception as e:
    print(tag,"failed",e); print(open(f.replace(".json",".err")).read()[-3000:])
PY
}
timeout 600 python bench.py --steps 3 --warmup 2 --e2e-steps 2 > gpurun_out/r2m_cfg3.json 2> gpurun_out/r2m_cfg3.err; summ "cfg3 filter" gpurun_out/r2m_cfg3.json
timeout 900 python bench.py --workload cfg5slab --steps 2 --warmup 1 --e2e-steps 1 --parity-windows 2 --parity-samples 100 > gpurun_out/r2m_cfg5slab.json 2> gpurun_out/r2m_cfg5slab.err; summ "cfg5slab filter" gpurun_out/r2m_cfg5slab.json
QMB_SLOW_NOFILTER=1 timeout 900 python bench.py --workload cfg5slab --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2m_cfg5slab_nofilter.json 2> gpurun_out/r2m_cfg5slab_nofilter.err; summ "cfg5slab nofilter" gpurun_out/r2m_cfg5slab_nofilter.json
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_scan -c 3 --csv --log-file gpurun_out/r2m_dram_cfg3.csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2m_ncu_cfg3.log 2>&1
tail -12 gpurun_out/r2m_dram_cfg3.csv
