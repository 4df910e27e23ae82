#!/bin/bash
# GPU visit for the 16-bit scan: full GPU suite, bench in both operand formats, ncu of k_scan.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/gpu_tests.log; tail -30 gpurun_out/gpu_tests.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
for prec in ${PRECS:-auto}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --precision $prec $BENCH_EXTRA > gpurun_out/bench_n1_$prec.json 2> gpurun_out/bench_n1_$prec.err
  python - <<PY
import json
j = json.load(open("gpurun_out/bench_n1_$prec.json"))
print("$prec value %.4g ms/step %.2f scan_ms %.2f e2e %.4g slow %s pipe %s" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["e2e"]["value"], j["slow_path_samples"], json.dumps(j["roofline"]["smem_pipe"])))
PY
  tail -3 gpurun_out/bench_n1_$prec.err
done
for form in ${FORMS:-1 3}; do
  QMB_Q16_FORM=$form timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_form$form.json 2> gpurun_out/bench_form$form.err
  python - <<PY
import json
j = json.load(open("gpurun_out/bench_form$form.json"))
print("q16 form $form: value %.4g scan_ms %.2f slow %s" % (j["value"], j["roofline"]["kernel_ms"], j["slow_path_samples"]))
PY
done
for extra in "$@"; do
  case $extra in
    ncu) timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/prof_scan_q16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_d.log 2>&1
         timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_d.log 2>&1;;
    cfg3) timeout 300 python bench.py --workload cfg3 --samples 36000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 1200 gpurun_out/bench_cfg3.json; tail -3 gpurun_out/bench_cfg3.err;;
  esac
done
ls -la gpurun_out | tail -12
