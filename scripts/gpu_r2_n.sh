#!/bin/bash
# Round 2, visit N (1 GPU): compact remainder tiles + unrolled k_slow -- parity tests and the cfg5 slab's tail.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ranges.py tests/test_gpu_sizes.py tests/test_gpu_multirank.py -q -m gpu -x 2>&1 > gpurun_out/r2n_gpu_tests.log; tail -5 gpurun_out/r2n_gpu_tests.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    line=[l for l in open(f).read().splitlines() if l.startswith("{")][-1]
    j=json.loads(line)
    print("%s value %.4g ms/step %.2f scan %.2f tail %.2f slow %s e2e %.4g parity %s span %s variant %s"%(tag,j["value"],j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["slow_path_samples"],j["e2e"]["value"],j.get("parity_vs_reference_sample",{}).get("pass"),j["config"]["max_tile_span"],j["config"]["variant"]))
    print("   scan", j.get("scan"))
except Exception as e:
    print(tag,"failed",e); print(open(f.replace(".json",".err")).read()[-3000:])
PY
}
timeout 900 python bench.py --workload cfg5slab --steps 2 --warmup 1 --e2e-steps 1 --parity-windows 2 --parity-samples 100 > gpurun_out/r2n_cfg5slab.json 2> gpurun_out/r2n_cfg5slab.err; summ "cfg5slab" gpurun_out/r2n_cfg5slab.json
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 --parity-windows 2 > gpurun_out/r2n_cfg2.json 2> gpurun_out/r2n_cfg2.err; summ "cfg2" gpurun_out/r2n_cfg2.json
