#!/bin/bash
# Round 2, visit Q (1 GPU): own FFT for the Hilbert envelopes (parity + timing), locate line diagnostics.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_reference_adapter.py tests/test_host_scan_loop.py -q -m gpu -x 2>&1 > gpurun_out/r2q_gpu_tests.log; tail -15 gpurun_out/r2q_gpu_tests.log
timeout 300 python - <<'PY'
import time, numpy as np
from quakemigrate_b200.core import Engine
eng = Engine(0)
rng = np.random.default_rng(0)
for C, n in ((48, 137460), (96, 409260)):
    wf = rng.normal(size=(C, n))
    dev = eng.upload_waveforms(wf)
    rows = [[c] for c in range(C)]
    for tr in ("energy", "env_squared"):
        ts = []
        for i in range(4):
            eng.sync(); t0 = time.perf_counter()
            o = eng.onsets_from_device_waveforms(dev, rows, 21, 201, position="classic", transform=tr)
            eng.sync(); ts.append(time.perf_counter() - t0); o.free()
        print("onset rows C=%d n=%d transform=%s: %.3f ms" % (C, n, tr, 1e3 * min(ts[1:])))
    dev.free()
eng.close()
PY
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 2 > gpurun_out/r2q_cfg4.json 2> gpurun_out/r2q_cfg4.err; tail -3 gpurun_out/r2q_cfg4.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2q_cfg4.json") if l.startswith("{")][-1])
print("cfg4 ms/step", j["ms_per_step"], "e2e", j["e2e"]["seconds_per_event"], "stages", {k:v for k,v in j["stages"].items() if k!="window_pass"}, "parity", j["parity_vs_reference_sample"]["pass"])
PY
