#!/bin/bash
# One GPU-box visit: full GPU test-suite, bench, optional extras.
set -x
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1800 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
for extra in "$@"; do
  case $extra in
    cfg3) python bench.py --workload cfg3 --samples 36000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 1500 gpurun_out/bench_cfg3.json; tail -3 gpurun_out/bench_cfg3.err;;
    ncu) ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
         ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/prof_scan python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1;;
    smallT) for T in 12000 2000; do python bench.py --samples $T --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('T', j['config']['n_samples'], 'value %.4g' % j['value'], 'ms/step %.3f' % j['ms_per_step'], 'scan_ms %.3f' % j['roofline']['kernel_ms'], 'e2e %.4g' % j['e2e']['value'], 'e2e_ms %.3f' % (j['config']['n_nodes']*j['config']['n_samples']/j['e2e']['value']*1e3))
"; done;;
    sanitize) timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "random or ties or window or legacy" 2>&1 | tail -15;;
  esac
done
ls -la gpurun_out | tail -12
