# -*- coding: utf-8 -*-
"""
File-level drop-in (tier T1): install libqmb200.so under the name the reference looks for
(``quakemigrate/core/src/qmlib<EXT_SUFFIX>``, core/libnames.py:35-38), load it the way the
reference does and drive it with the reference's own ctypes argtypes (core/lib.py:27-49,128,173,
211,249) -- nothing from quakemigrate_b200's Python layer is involved.
"""

import ctypes
import pathlib
import shutil
import sysconfig

import numpy as np
import numpy.ctypeslib as clib
import pytest

from oracle import qm_oracle as O
from quakemigrate_b200 import _lib, synthetic as syn

pytestmark = pytest.mark.gpu


def _load_like_the_reference(tmp_path):
    src_dir = tmp_path / "quakemigrate" / "core" / "src"
    src_dir.mkdir(parents=True)
    lib = (src_dir / "qmlib").with_suffix(sysconfig.get_config_var("EXT_SUFFIX"))
    shutil.copy(_lib.LIB_PATH, lib)
    return ctypes.CDLL(str(lib))


def test_reference_loader_and_argtypes(tmp_path):
    qmlib = _load_like_the_reference(tmp_path)
    c_int32, c_int64 = ctypes.c_int32, ctypes.c_int64
    c_dPt = clib.ndpointer(dtype=np.double, flags="C_CONTIGUOUS")
    c_i32Pt = clib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    c_i64Pt = clib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
    header_t = np.dtype([("n", c_int32), ("nsta", c_int32), ("nlta", c_int32)], align=True)
    header_pt = clib.ndpointer(header_t, flags="C_CONTIGUOUS")
    qmlib.migrate.argtypes = [c_dPt, c_i32Pt, c_dPt, c_int32, c_int32, c_int32, c_int32, c_int32, c_int64, c_int64]
    qmlib.find_max_coa.argtypes = [c_dPt, c_dPt, c_dPt, c_i64Pt, c_int32, c_int64, c_int64]
    qmlib.overlapping_sta_lta.argtypes = [c_dPt, header_pt, c_dPt]

    grid, S, T, fsmp, lsmp = (6, 7, 5), 5, 150, 6, 30
    onsets, tau = syn.random_problem(31, grid, S, T, fsmp, lsmp, neg_frac=0.05)
    log_on = np.log(np.clip(onsets, 0.01, np.inf))             # lib.py:93-94
    n_nodes = int(np.prod(grid))
    map4d = np.zeros(grid + (T,), dtype=np.double)              # lib.py:101
    qmlib.migrate(log_on, tau, map4d, c_int32(fsmp), c_int32(lsmp), c_int32(T), c_int32(S), c_int32(S),
                  c_int64(n_nodes), c_int64(4))
    want = O.migrate(onsets, tau, fsmp, lsmp, S, threads=2)
    np.testing.assert_allclose(map4d, want, rtol=1e-5)

    # accumulate-into semantics of migratelib.c:57: a second call adds onto the (non-zero) array
    twice = np.log(want)  # log(exp(x/avail)) = x/avail ... feed something known instead:
    pre = np.full_like(want, 0.25)
    acc = pre.copy()
    qmlib.migrate(log_on, tau, acc, c_int32(fsmp), c_int32(lsmp), c_int32(T), c_int32(S), c_int32(S),
                  c_int64(n_nodes), c_int64(1))
    np.testing.assert_allclose(acc, want * np.exp(0.25 / S), rtol=1e-5)
    del twice

    max_coa, max_norm, idx = np.zeros(T), np.zeros(T), np.zeros(T, dtype=np.int64)
    qmlib.find_max_coa(want, max_coa, max_norm, idx, c_int32(T), c_int64(n_nodes), c_int64(4))
    r = O.find_max_coa(want, 2)
    assert np.array_equal(idx, r[2]) and np.array_equal(max_coa, r[0])
    np.testing.assert_allclose(max_norm, r[1], rtol=1e-13)

    sig = np.arange(6, dtype=np.double)
    head = np.empty(1, dtype=header_t)
    head[:] = (6, 2, 3)
    onset = np.ones(6)
    qmlib.overlapping_sta_lta(sig, head, onset)
    assert (onset == np.array([1.0, 1.0, 1.5, 1.25, 21. / 18, 27. / 24])).all()   # tests/test_onsets.py:28-33


def test_config1_full_size_through_the_reference_symbols():
    """BASELINE config 1 at FULL size: the real 71x64x57 Iceland table (rebuilt from the committed
    facts, identical to examples/benchmarks/Icequake_Iceland/example.LUT -- see
    tests/test_iceland_locate.py), the availability pattern of the example run (24 rows), 250 Hz,
    T = 625, fsmp = 413, lsmp = 1000, through quakemigrate_b200.core.migrate / find_max_coa (the five
    T1 symbols) against the compiled reference."""
    import pandas as pd
    import iceland_example as ice
    from quakemigrate_b200 import core

    lut = ice.build_lut()
    av = pd.read_csv(pathlib.Path(__file__).parent / "golden" / "golden_2014_180_StationAvailability.csv")
    availability = {k: int(av.iloc[0][k]) for k in av.columns[1:]}
    tau = lut.serve_traveltimes(250, availability)
    assert tau.shape == (71, 64, 57, 24) and tau.max() == 382
    S, T, fsmp, lsmp = 24, 625, 413, 1000
    rng = np.random.default_rng(180)
    onsets = np.clip(rng.gamma(2.0, 0.55, size=(S, fsmp + T + lsmp)), 0.4, np.inf)
    node = (40, 31, 20)
    for s in range(S):
        a = fsmp + 300 + int(tau[node][s])
        onsets[s, a - 6:a + 7] += 8.0 * np.hanning(13)
    map4d = core.migrate(onsets, tau, fsmp, lsmp, S, 4)
    want = O.migrate(onsets, tau, fsmp, lsmp, S, threads=O.default_threads())
    np.testing.assert_allclose(map4d, want, rtol=1e-5)
    got = core.find_max_coa(want, 4)
    ref = O.find_max_coa(want, threads=O.default_threads())
    assert np.array_equal(got[2], ref[2]) and np.array_equal(got[0], ref[0])
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-13)
    assert ref[2][300] == np.ravel_multi_index(node, tau.shape[:-1])
    # and the fused T2 path on the same inputs: indices exact against the reference
    with core.Engine(0) as eng:
        dl = eng.upload_lut(tau)
        mx, nm, ix = eng.detect(dl, onsets, fsmp, lsmp, S)
        assert np.array_equal(ix, ref[2])
        np.testing.assert_allclose(mx, ref[0], rtol=1e-9)
        np.testing.assert_allclose(nm, ref[1], rtol=1e-5)


def test_failures_raise_instead_of_aborting():
    """The C symbols are void (qmlib.h:28-44); a failure inside them is recorded and the Python layer
    raises -- the interpreter survives."""
    from quakemigrate_b200 import core
    grid, S, T, fsmp, lsmp = (5, 4, 3), 4, 120, 5, 20
    onsets, tau = syn.random_problem(3, grid, S, T, fsmp, lsmp)
    tau[2, 1, 1, 0] = lsmp + 7                     # beyond the post-pad: the reference reads out of bounds
    with pytest.raises(ValueError):
        core.migrate(onsets, tau, fsmp, lsmp, S, 1)
    with pytest.raises(ValueError):
        core.overlapping_sta_lta(np.ones(50), 9, 4)  # nsta > nlta
    # straight through the C ABI (what an unmodified quakemigrate.core would do): no abort, status set
    lib = _lib.load()
    log_on = np.log(np.clip(onsets, 0.01, np.inf))
    map4d = np.zeros(grid + (T,))
    lib.migrate(log_on, np.ascontiguousarray(tau), map4d, fsmp, lsmp, T, S, S, int(np.prod(grid)), 1)
    assert lib.qmb_legacy_status() == 2            # QMB_ERR_ARG
    assert b"travel time" in lib.qmb_last_error()
    assert lib.qmb_legacy_status() == 0            # cleared by the read
    assert not map4d.any()                          # output untouched
