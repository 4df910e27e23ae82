#!/bin/bash
# 2-GPU visit after the 16-bit scan: NCCL parity test + precision tests + N=2 bench.
set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_precision.py -q 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_multi_n2.json 2> gpurun_out/bench_multi_n2.err
python - <<'PY'
import json
for line in open("gpurun_out/bench_multi_n2.json"):
    try:
        j = json.loads(line)
    except Exception:
        continue
    print("n_gpus", j["n_gpus"], "value %.4g" % j["value"], "ms/step %.2f" % j["ms_per_step"], "e2e %.4g" % j["e2e"]["value"], j["config"]["scan_operands"], "slow", j["slow_path_samples"])
PY
tail -3 gpurun_out/bench_multi_n2.err
