# -*- coding: utf-8 -*-
"""
Generate the committed golden vectors from the REFERENCE ITSELF.

Run in the dev container (needs /root/reference): the reference C sources are compiled unmodified
with the reference flags into oracle/_ref/qmlib.so (oracle/Makefile) and driven through the same
ctypes argtypes / pre-processing as quakemigrate/core/lib.py.  Outputs:

  golden_small.npz     random stress problem (negative tau, clip-floor ties) + STA/LTA vectors
  golden_iceland.npz   the bundled example LUT (examples/benchmarks/Icequake_Iceland/example.LUT:
                       71x64x57 nodes at 25 m, 13 stations x {P,S}), served at 250 Hz for the
                       availability pattern of 2014_180_StationAvailability.csv row 0 (24 rows),
                       decimated by 3 nodes per axis to keep the fixture small, with synthetic
                       onsets (S=24, T=625, fsmp=413, lsmp=1000 -- BASELINE config 1 shape).

    python tests/golden/make_golden.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import qm_oracle as O  # noqa: E402
from quakemigrate_b200 import synthetic as syn  # noqa: E402
from quakemigrate_b200.lut import read_reference_lut  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path("/root/reference")


def small():
    grid, S, T, fsmp, lsmp = (7, 6, 5), 9, 160, 11, 45
    onsets, tau = syn.random_problem(2024, grid, S, T, fsmp, lsmp, neg_frac=0.05, floor_frac=0.5)
    m4 = O.migrate(onsets, tau, fsmp, lsmp, S, threads=4, kind="ref")
    mx, nm, ix = O.find_max_coa(m4, threads=4, kind="ref")
    rng = np.random.default_rng(7)
    sig = rng.standard_normal(1500) ** 2
    sig[700:720] += 40.0
    np.savez_compressed(
        HERE / "golden_small.npz", onsets=onsets, tau=tau, fsmp=fsmp, lsmp=lsmp, available=S,
        map4d=m4.astype(np.float32), max_coa=mx, max_norm_coa=nm, max_coa_idx=ix,
        map4d_checksum=np.array([m4.sum(), (m4 ** 2).sum()]),
        signal=sig, nsta=21, nlta=101,
        classic=O.overlapping_sta_lta(sig, 21, 101, kind="ref"),
        centred=O.centred_sta_lta(sig, 21, 101, kind="ref"),
        recursive=O.recursive_sta_lta(sig, 21, 101, kind="ref"))


def iceland():
    import pandas as pd
    bench = REF / "examples" / "benchmarks" / "Icequake_Iceland"
    lut = read_reference_lut(bench / "example.LUT")
    av = pd.read_csv(bench / "2014_180_StationAvailability.csv")
    availability = {k: int(av.iloc[0][k]) for k in av.columns[1:]}
    full = lut.serve_traveltimes(250, availability)             # (71, 64, 57, 24) int32
    assert full.shape == (71, 64, 57, 24) and full.max() == 382
    dec = np.ascontiguousarray(full[::3, ::3, ::3])
    S, T, fsmp, lsmp = 24, 625, 413, 1000
    rng = np.random.default_rng(180)
    onsets = np.clip(rng.gamma(2.0, 0.55, size=(S, fsmp + T + lsmp)), 0.4, np.inf)
    # one planted event: raise the onsets along the moveout of a node
    node = (12, 9, 7)
    for s in range(S):
        a = fsmp + 300 + int(dec[node][s])
        onsets[s, a - 6:a + 7] += 8.0 * np.hanning(13)
    mx, nm, ix = O.find_max_coa(O.migrate(onsets, dec, fsmp, lsmp, S, threads=8, kind="ref"), threads=8, kind="ref")
    assert ix[300] == np.ravel_multi_index(node, dec.shape[:-1])
    np.savez_compressed(HERE / "golden_iceland.npz", tau=dec.astype(np.int16), onsets=onsets, fsmp=fsmp,
                        lsmp=lsmp, available=S, availability_keys=np.array(list(availability)),
                        availability_vals=np.array(list(availability.values())),
                        max_coa=mx, max_norm_coa=nm, max_coa_idx=ix, planted_node=np.array(node),
                        full_lut_checksum=np.array([int(full.astype(np.int64).sum()), int(full.max())]))


if __name__ == "__main__":
    assert O.have("ref"), "build oracle/_ref first: make -C oracle"
    small()
    iceland()
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)
