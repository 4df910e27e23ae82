#!/usr/bin/env python
"""
Time the locate path on BASELINE config 4 (151^3 nodes, 48 onset rows, 30000 samples, centred
STA/LTA): max-coalescence traces over the whole span, marginal window around the peak
(io/event.py:421-439) re-migrated on the device, location fits (scan.py:696-733).

    python scripts/locate_bench.py [--samples 30000] [--grid 151 151 151]
Prints one JSON line.
"""
import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from quakemigrate_b200 import synthetic as syn  # noqa: E402
from quakemigrate_b200.core import Engine  # noqa: E402
from quakemigrate_b200.locate import calculate_location  # noqa: E402
from quakemigrate_b200.lut import Grid3D  # noqa: E402
from quakemigrate_b200.scan import marginal_window_indices  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=30000)
    ap.add_argument("--grid", type=int, nargs=3, default=None)
    ap.add_argument("--marginal-window", type=float, default=2.0)
    a = ap.parse_args()
    cfg = syn.CONFIGS["cfg4"].scaled(grid=a.grid, n_samples=a.samples, name="cfg4")
    t0 = time.perf_counter()
    prob = syn.make_problem(cfg)
    t_gen = time.perf_counter() - t0
    eng = Engine(0)
    t0 = time.perf_counter()
    lut = eng.upload_lut(prob.traveltimes)
    eng.sync()
    t_lut = time.perf_counter() - t0
    grid = Grid3D(cfg.grid, [cfg.node_spacing_km] * 3, [0.0, 0.0, 0.0])
    out = {}
    for rep in range(3):
        t0 = time.perf_counter()
        dev = eng.onsets_from_waveforms(prob.waveforms, prob.row_channels, prob.nsta, prob.nlta,
                                        position=cfg.position)
        t1 = time.perf_counter()
        ms = eng.detect_device(lut, dev, prob.fsmp, prob.lsmp, cfg.n_rows)
        mx, nm, ix = eng.detect_fetch(cfg.n_samples)
        t2 = time.perf_counter()
        i0, i1, imax = marginal_window_indices(mx, a.marginal_window, cfg.sampling_rate)
        coa = eng.locate_window(lut, dev, prob.fsmp, prob.lsmp, cfg.n_rows, i0, i1)
        t3 = time.perf_counter()
        loc = calculate_location(coa, grid, eng)
        t4 = time.perf_counter()
        dev.free()
        out = {"onsets_s": t1 - t0, "detect_pass_s": t2 - t1, "scan_kernel_ms": ms["scan_ms"],
               "window_pass_s": t3 - t2, "fits_s": t4 - t3, "total_s": t4 - t0,
               "window": [i0, i1], "peak_sample": imax, "slow_path_samples": ms["slow_samples"]}
    node = np.unravel_index(ix[imax], cfg.grid)
    line = {"what": "locate path, BASELINE config 4", "grid": list(cfg.grid), "n_rows": cfg.n_rows,
            "n_samples": cfg.n_samples, "position": cfg.position, "lut_upload_s": t_lut,
            "host_problem_generation_s": t_gen, **out,
            "detect_node_samples_per_s": cfg.n_nodes * cfg.n_samples / out["detect_pass_s"],
            "peak_node": [int(v) for v in node], "planted_node": [int(v) for v in np.unravel_index(prob.events[0][0], cfg.grid)],
            "planted_origin_sample": int(prob.events[0][1]),
            "gaussian_location_km": [float(v) for v in loc["gaussian"][0]],
            "gaussian_uncertainty_km": [float(v) for v in loc["gaussian"][1]],
            "covariance_location_km": [float(v) for v in loc["covariance"][0]],
            "spline_location_km": [float(v) for v in loc["spline"]]}
    print(json.dumps(line))
    eng.close()


if __name__ == "__main__":
    main()
