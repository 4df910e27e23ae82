#!/bin/bash
# Round 2, visit X (2 GPUs): multi-rank parity tests and a short cfg3 N=2 line on the final code.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --e2e-steps 2 --parity-windows 2 > gpurun_out/r2x_cfg3_n2.json 2> gpurun_out/r2x_cfg3_n2.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2x_cfg3_n2.json") if l.startswith("{")][-1])
    print("cfg3 N=2 value %.4g ms/step %.2f tail %.2f slow %s e2e %.4g parity %s per_rank %s"%(j["value"],j["ms_per_step"],j["tail_ms_per_step"],j["slow_path_samples"],j["e2e"]["value"],j["parity_vs_reference_sample"]["pass"],j["per_rank"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2x_cfg3_n2.err").read()[-3000:])
PY
