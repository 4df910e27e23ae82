# -*- coding: utf-8 -*-
"""
Golden vectors for the waveform pre-processing stage, generated with the
libraries the reference itself calls through obspy (scipy.signal: detrend, iirfilter, zpk2sos,
sosfilt) via oracle/preprocess_oracle.py.  Run in the dev container:

    python tests/golden/make_golden_preprocess.py  ->  tests/golden/golden_preprocess.npz

Contents: three synthetic traces (noise + offset + trend + a transient), their pre_process output
for the filters of the reference's examples ([10, 124, 4] at 250 Hz,
examples/Icequake_Iceland/iceland_detect.py:45-47; [2, 16, 2] at 50 Hz,
examples/Volcanotectonic_Iceland/dike_intrusion_detect.py:47-50) and a narrow low-frequency band,
util.decimate 100 Hz -> 50 Hz, detrend + taper alone, and the filter sections.
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import preprocess_oracle as PO  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent

if __name__ == "__main__":
    rng = np.random.default_rng(424242)
    n = 4000
    t = np.arange(n)
    raw = rng.standard_normal((3, n)) * np.array([[120.0], [3.0], [4.0e4]])
    raw += np.array([[1.5e4], [-20.0], [3.0e6]]) + np.array([[0.3], [-0.002], [55.0]]) * t
    raw[0, 2000:2100] += 3000.0 * np.sin(2 * np.pi * 8.0 * t[:100] / 50.0) * np.hanning(100)
    raw = np.round(raw)  # archive data are integer counts
    out = {"raw": raw}
    for name, df, filt in (("vt", 50.0, [2.0, 16.0, 2]), ("ice", 250.0, [10.0, 124.0, 4]),
                           ("low", 100.0, [0.1, 0.5, 3])):
        sos = PO.butter_sos("bandpass", filt[:2], df, filt[2])
        out[f"sos_{name}"] = sos
        out[f"df_{name}"] = df
        out[f"filt_{name}"] = np.array(filt, dtype=float)
        out[f"pre_{name}"] = np.stack([PO.pre_process(x, df, filt) for x in raw])
    out["sos_decimate"] = PO.butter_sos("lowpass", 50.0 / 2.000001, 100.0, 2)
    out["decimated"] = np.stack([PO.decimate(x, 100.0, 50.0) for x in raw])
    out["detrend_taper"] = np.stack([PO.detrend_taper(x) for x in raw])
    np.savez_compressed(HERE / "golden_preprocess.npz", **out)
    print("wrote", HERE / "golden_preprocess.npz")
