#!/bin/bash
# Round 2, visit Y (1 GPU): per-kernel time / DRAM traffic of the own envelope transforms (cfg3's trace block).
set -x
mkdir -p gpurun_out
cat > /tmp/env_probe.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from quakemigrate_b200.core import Engine
eng = Engine(0)
rng = np.random.default_rng(0)
wf = rng.normal(size=(96, 409260))
dev = eng.upload_waveforms(wf)
rows = [[c] for c in range(96)]
for i in range(2):
    o = eng.onsets_from_device_waveforms(dev, rows, 21, 201, position="classic", transform="env_squared")
    eng.sync(); o.free()
dev.free(); eng.close()
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum --clock-control none -k regex:'k_fft_pass|k_chirp|k_analytic|k_envelope' -c 80 --csv --log-file gpurun_out/r2y_envelope_kernels.csv python /tmp/env_probe.py > gpurun_out/r2y_ncu.log 2>&1
tail -3 gpurun_out/r2y_ncu.log; wc -l gpurun_out/r2y_envelope_kernels.csv
