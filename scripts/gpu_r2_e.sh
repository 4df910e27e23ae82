#!/bin/bash
# Round 2, visit E (2 GPUs): sharded parity on hardware + cfg3 at N=2 with the parity block.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_iceland_locate.py tests/test_gpu_dropin.py tests/test_locate_stage.py tests/test_gpu_scan_api.py -q -m gpu 2>&1 > gpurun_out/r2e_gpu_tests.log; tail -12 gpurun_out/r2e_gpu_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2e_cfg3_n2.json 2> gpurun_out/r2e_cfg3_n2.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r2e_cfg3_n2.json"))
    print("cfg3 N=2 value %.4g ms/step %.2f scan %.2f tail %.2f slow %s parity %s e2e %.4g"%(j["value"],j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["slow_path_samples"],j["parity_vs_reference_sample"],j["e2e"]["value"]))
except Exception as e:
    print("N=2 failed", e); print(open("gpurun_out/r2e_cfg3_n2.err").read()[-3000:])
PY
