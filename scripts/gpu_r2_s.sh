#!/bin/bash
# Round 2, visit S (1 GPU): tile-geometry sweep of the fused pre-processing kernel (samples per thread x threads per tile).
set -x
mkdir -p gpurun_out
run() { # tag, lib
  for wl in cfg2 cfg3; do
    QMB_LIB=$2 timeout 200 python bench.py --pre-only --workload $wl --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r2s_$1_$wl.json 2> gpurun_out/r2s_$1_$wl.err
    python - "$1 $wl" gpurun_out/r2s_$1_$wl.json <<'PY'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(sys.argv[1], "kernel_ms %.4f frac %.4f"%(j["kernel_ms"], j["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e); print(open(sys.argv[2].replace(".json",".err")).read()[-1500:])
PY
  done
}
run base quakemigrate_b200/libqmb200.so
for v in 16_256 16_512 8_512 32_128 64_128; do run $v experiments/libqmb200_pp_$v.so; done
