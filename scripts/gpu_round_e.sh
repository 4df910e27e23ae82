#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/gpu_tests.log; tail -4 gpurun_out/gpu_tests.log
for prec in auto fp32; do
timeout 300 python bench.py --steps 5 --warmup 3 --precision $prec $( [ $prec = fp32 ] && echo --no-cpu-baseline ) > gpurun_out/bench_n1_$prec.json 2> gpurun_out/bench_n1_$prec.err
python - <<PY
import json
j = json.load(open("gpurun_out/bench_n1_$prec.json"))
print("$prec value %.4g ms/step %.2f scan_ms %.2f e2e %.4g slow %s" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["e2e"]["value"], j["slow_path_samples"]))
PY
done
