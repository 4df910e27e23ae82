#!/bin/bash
# Multi-GPU visit: NCCL parity test + weak-scaling bench at N GPUs (and N=1 on the same box).
N=${1:-2}
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_scan_api.py -q 2>&1 | tail -8
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_multi_n1.json 2> gpurun_out/bench_multi_n1.err
tail -c 600 gpurun_out/bench_multi_n1.json; tail -3 gpurun_out/bench_multi_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_multi_n$n.json 2> gpurun_out/bench_multi_n$n.err
    tail -c 900 gpurun_out/bench_multi_n$n.json; tail -5 gpurun_out/bench_multi_n$n.err
  fi
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_multi_n*.json')):
    for line in open(f):
        try: j = json.loads(line)
        except Exception: continue
        print(f, 'n_gpus', j['n_gpus'], 'value %.4g' % j['value'], 'ms/step %.1f' % j['ms_per_step'], 'e2e %.4g' % j['e2e']['value'])
PY
