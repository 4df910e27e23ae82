#!/bin/bash
# Round 2, visit O (8 GPUs): final scaling lines -- cfg3 grid-sharded over 8 and 4 GPUs, BASELINE cfg5 at full size over 8.
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
summ() { python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    line=[l for l in open(f).read().splitlines() if l.startswith("{")][-1]
    j=json.loads(line)
    print("%s value %.4g ms/step %.2f scan %.2f tail %.2f slow %s e2e %.4g"%(tag,j["value"],j["ms_per_step"],j["roofline"]["kernel_ms"],j["tail_ms_per_step"],j["slow_path_samples"],j["e2e"]["value"]))
    print("   parity", j.get("parity_vs_reference_sample"))
    print("   per_rank", j.get("per_rank"))
    print("   scan", j.get("scan"))
except Exception as e:
    print(tag,"failed",e); print(open(f.replace(".json",".err")).read()[-3000:])
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2o_cfg3_n8.json 2> gpurun_out/r2o_cfg3_n8.err; summ "cfg3 N=8" gpurun_out/r2o_cfg3_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 5 --warmup 3 --parity-windows 2 > gpurun_out/r2o_cfg3_n4.json 2> gpurun_out/r2o_cfg3_n4.err; summ "cfg3 N=4" gpurun_out/r2o_cfg3_n4.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 --parity-windows 2 --parity-samples 100 > gpurun_out/r2o_cfg5_n8.json 2> gpurun_out/r2o_cfg5_n8.err; summ "cfg5 N=8" gpurun_out/r2o_cfg5_n8.json
