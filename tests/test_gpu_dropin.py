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
