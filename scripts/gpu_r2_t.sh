#!/bin/bash
# Round 2, visit T (1 GPU): where the time before the scan goes at cfg3 size (upload / filter / onset rows).
set -x
timeout 300 python - <<'PY'
import time, numpy as np
from quakemigrate_b200.core import Engine
from quakemigrate_b200 import preprocess as pp
eng = Engine(0)
rng = np.random.default_rng(0)
for C, n, S in ((48, 137460, 32), (96, 409260, 64)):
    wf = rng.normal(size=(C, n))
    sos = pp.butter_sos("bandpass", (2.0, 16.0), 100, 2)
    rows = [[c] for c in range(S // 2)] + [[S // 2 + 2 * c % (C - S // 2), S // 2 + (2 * c + 1) % (C - S // 2)] for c in range(S // 2)]
    for it in range(3):
        eng.sync(); t0 = time.perf_counter()
        raw = eng.upload_waveforms(wf); eng.sync(); t1 = time.perf_counter()
        flt = eng.filter_waveforms(raw, sos); eng.sync(); t2 = time.perf_counter()
        dev = eng.onsets_from_device_waveforms(flt, rows, 21, 201, position="classic"); eng.sync(); t3 = time.perf_counter()
        raw.free(); flt.free(); dev.free(); eng.sync(); t4 = time.perf_counter()
        print("C=%d n=%d: upload %.2f ms, filter %.2f ms, onset rows %.2f ms, frees %.2f ms" % (C, n, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)))
    Engine.pin(wf)
    for it in range(2):
        eng.sync(); t0 = time.perf_counter()
        raw = eng.upload_waveforms(wf); eng.sync(); t1 = time.perf_counter()
        raw.free()
        print("   pinned upload %.2f ms" % (1e3*(t1-t0)))
    Engine.unpin(wf)
eng.close()
PY
