#!/bin/bash
# parity for every variant, then perf of the variants given as arguments (default: all)
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
for v in ${@:-0 1 2 3 4 5 6 7}; do python bench.py --steps 3 --warmup 2 --no-cpu-baseline --variant $v 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: j = json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    print('variant', j['config']['variant'], 'value %.4g' % j['value'], 'scan_ms %.2f' % j['roofline']['kernel_ms'], 'smem_frac %.4f' % j['roofline']['smem_pipe']['frac'], 'span', j['config']['max_tile_span'], 'e2e %.4g' % j['e2e']['value'])
"; done
