#!/bin/bash
# perf sweep: rows-per-stage x kernel variant (cfg2, full T)
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
for cr in 16 8 32; do for v in 4 2; do
QMB_CHUNK_ROWS=$cr python bench.py --steps 3 --warmup 2 --no-cpu-baseline --variant $v 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: j = json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    print('chunk_rows $cr variant', j['config']['variant'], 'value %.4g' % j['value'], 'scan_ms %.2f' % j['roofline']['kernel_ms'], 'smem_frac %.4f' % j['roofline']['smem_pipe']['frac'], 'e2e %.4g' % j['e2e']['value'])
"; done; done
QMB_CHUNK_ROWS=8 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
QMB_CHUNK_ROWS=32 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
