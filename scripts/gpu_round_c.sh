#!/bin/bash
# Short GPU visit: pre-processing tests + bench + launch list.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_preprocess.py -q -m gpu 2>&1 | tail -40 > gpurun_out/pre_tests.log; tail -12 gpurun_out/pre_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_n1.json"))
print("value %.4g e2e %.4g" % (j["value"], j["e2e"]["value"]))
print(json.dumps(j["stages"], indent=1))
PY
tail -3 gpurun_out/bench_n1.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c.csv -k regex:"k_sos|k_pre" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_c.log 2>&1
for extra in "$@"; do
  case $extra in
    cfg3pre) timeout 200 python scripts/preprocess_bench.py > gpurun_out/preprocess_cfg3.json 2> gpurun_out/preprocess_cfg3.err; cat gpurun_out/preprocess_cfg3.json; tail -3 gpurun_out/preprocess_cfg3.err;;
    full) timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_sos_state|k_sos_apply" -c 4 -o gpurun_out/prof_sos python bench.py --steps 1 --warmup 1 --samples 12000 --no-cpu-baseline --pre-only > gpurun_out/ncu_full_c.log 2>&1;;
  esac
done
ls -la gpurun_out | tail -12
