#!/usr/bin/env python
"""
Pre-processing stage at BASELINE config 3's size: 96 traces (32 stations x 3 components) x
(360000 + pads) samples, pre_process [2, 16, 2] at 100 Hz.  Prints one JSON object.
"""
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from quakemigrate_b200 import preprocess as pp  # noqa: E402
from quakemigrate_b200.core import Engine  # noqa: E402

if __name__ == "__main__":
    c, n, sr = 96, 381_000, 100.0
    rng = np.random.default_rng(0)
    wf = rng.standard_normal((c, n)) * 100.0 + 5000.0
    sos = pp.butter_sos("bandpass", (2.0, 16.0), sr, 2)
    peak = 6547.2
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    with Engine(0) as eng:
        raw = eng.upload_waveforms(wf)
        out = eng.alloc_waveforms(c, n)
        for _ in range(3):
            eng.filter_waveforms(raw, sos, out=out)
        ms = []
        for _ in range(10):
            eng.flush_l2()
            _, t = eng.filter_waveforms(raw, sos, out=out, timing=True)
            ms.append(t)
        Engine.pin(wf)
        res = np.empty_like(wf)
        Engine.pin(res)
        eng.sync()
        t0 = time.perf_counter()
        for _ in range(3):
            eng._lib.qmb_preprocess(eng._h, wf, c, n, sos.ctypes.data, 2, 1, 0.05, 1, 1, res)
        e2e = (time.perf_counter() - t0) / 3
    t_ms = float(np.median(ms))
    print(json.dumps({
        "stage": "pre_process [2, 16, 2] @ 100 Hz", "channels": c, "samples_per_channel": n,
        "kernel_ms": t_ms, "kernel_ms_all": ms, "channel_samples_per_s": c * n / (t_ms * 1e-3),
        "algorithmic_GBps": 16.0 * c * n / (t_ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
        "roofline_frac": 16.0 * c * n / (t_ms * 1e-3) / 1e9 / peak,
        "e2e_ms_host_buffers": e2e * 1e3, "e2e_channel_samples_per_s": c * n / e2e,
        "h2d_bytes": int(wf.nbytes), "d2h_bytes": int(res.nbytes)}))
