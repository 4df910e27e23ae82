#!/usr/bin/env python
"""Sweep the shared-memory pipe micro-benchmark: vector width x warps per SM (B200)."""
import ctypes, pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from quakemigrate_b200 import _lib
from quakemigrate_b200.core import Engine

eng = Engine(0)
lib = _lib.load()
nominal = 148 * 32 * 1.965e9
print("| load | warps/SM | gather-adds/s | of nominal 32/clk/SM @1965 MHz |\n|---|---:|---:|---:|")
for vec in (1, 2, 4):
    for warps in (4, 8, 12, 13, 16, 24, 32):
        out = ctypes.c_double(0)
        _lib.check(lib.qmb_microbench_smem_pipe_ex(eng._h, 40000, vec, warps, ctypes.byref(out)))
        print(f"| LDS.{32*vec} | {warps} | {out.value:.4g} | {out.value / nominal:.4f} |")
print("k_scan-shaped (LDS.32, 13 warps, 32 loads/iter): %.4g" % eng.smem_pipe_peak(40000))
