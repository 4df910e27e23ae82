#!/bin/bash
# Round 2, visit P (1 GPU): device evaluation of the locate spline, timestep-chain stage of the bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_locate_stage.py tests/test_iceland_locate.py tests/test_host_scan_loop.py tests/test_gpu_sizes.py -q -m gpu -x 2>&1 > gpurun_out/r2p_gpu_tests.log; tail -5 gpurun_out/r2p_gpu_tests.log
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 2 > gpurun_out/r2p_cfg4.json 2> gpurun_out/r2p_cfg4.err; tail -3 gpurun_out/r2p_cfg4.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2p_cfg4.json") if l.startswith("{")][-1])
print("cfg4 ms/step", j["ms_per_step"], "e2e", j["e2e"], "stages", {k:v for k,v in j["stages"].items() if k.endswith("_ms") or k.startswith("location")}, "parity", j["parity_vs_reference_sample"]["pass"])
PY
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 --parity-windows 2 > gpurun_out/r2p_cfg2.json 2> gpurun_out/r2p_cfg2.err; tail -3 gpurun_out/r2p_cfg2.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2p_cfg2.json") if l.startswith("{")][-1])
print("cfg2 ms/step", j["ms_per_step"], "chain", j["stages"]["timestep_chain"])
PY
