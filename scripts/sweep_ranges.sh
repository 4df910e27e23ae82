#!/bin/bash
for cr in 32 16; do for R in 4 8 11 16 22 41; do
QMB_CHUNK_ROWS=$cr QMB_NODE_RANGES=$R python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: j = json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    print('chunk_rows $cr R $R value %.4g' % j['value'], 'scan_ms %.2f' % j['roofline']['kernel_ms'], 'total_ms %.2f' % j['ms_per_step'], 'smem_frac %.4f' % j['roofline']['smem_pipe']['frac'])
"; done; done
