#!/bin/bash
# Round 2, visit W (1 GPU): warp-shuffle scan in the fused pre-processing kernel -- parity tests + timing.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_reference_adapter.py tests/test_iceland_locate.py -q -m gpu -x 2>&1 > gpurun_out/r2w_gpu_tests.log; tail -5 gpurun_out/r2w_gpu_tests.log
for wl in cfg2 cfg3; do
  timeout 200 python bench.py --pre-only --workload $wl --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2w_pre_$wl.json 2> gpurun_out/r2w_pre_$wl.err
  python - "$wl" gpurun_out/r2w_pre_$wl.json <<'PY'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(sys.argv[1], "kernel_ms %.4f frac %.4f"%(j["kernel_ms"], j["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open(sys.argv[2].replace(".json",".err")).read()[-1500:])
PY
done
