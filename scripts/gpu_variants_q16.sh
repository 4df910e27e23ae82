#!/bin/bash
# q16 scan with more consumer warps: variants 8 / 9 (15 warps, 60-node tiles), 0 (16 warps, 96 regs)
mkdir -p gpurun_out
for v in ${VARS:-7 8 9 0}; do
  timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --variant $v > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_var$v.json"))
    print("variant $v: value %.4g scan_ms %.2f slow %s %s" % (j["value"], j["roofline"]["kernel_ms"], j["slow_path_samples"], j["config"]["scan_operands"]))
except Exception as e:
    print("variant $v failed", e)
PY
  tail -2 gpurun_out/bench_var$v.err
done
