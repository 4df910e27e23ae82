# -*- coding: utf-8 -*-
"""
Golden vectors produced by the reference C code itself (tests/golden/make_golden.py) pin
(a) the oracle port on CPU and (b) the CUDA path on the GPU, also where /root/reference and
oracle/_ref are absent.
"""

import pathlib

import numpy as np
import pytest

from oracle import qm_oracle as O

G = pathlib.Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def small():
    return np.load(G / "golden_small.npz")


@pytest.fixture(scope="module")
def iceland():
    return np.load(G / "golden_iceland.npz")


# ---- CPU: oracle port vs reference outputs --------------------------------------------------------
def test_port_matches_golden_small(small):
    g = small
    m4 = O.migrate(g["onsets"], g["tau"], int(g["fsmp"]), int(g["lsmp"]), int(g["available"]), threads=2)
    np.testing.assert_allclose(m4, g["map4d"], rtol=2e-7)            # stored as float32
    np.testing.assert_allclose([m4.sum(), (m4 ** 2).sum()], g["map4d_checksum"], rtol=1e-13)
    mx, nm, ix = O.find_max_coa(m4, 2)
    assert np.array_equal(ix, g["max_coa_idx"])
    np.testing.assert_allclose(mx, g["max_coa"], rtol=1e-14)
    np.testing.assert_allclose(nm, g["max_norm_coa"], rtol=1e-13)


def test_port_stalta_matches_golden(small):
    g = small
    nsta, nlta = int(g["nsta"]), int(g["nlta"])
    np.testing.assert_allclose(O.overlapping_sta_lta(g["signal"], nsta, nlta), g["classic"], rtol=1e-13)
    np.testing.assert_allclose(O.centred_sta_lta(g["signal"], nsta, nlta), g["centred"], rtol=1e-13)
    np.testing.assert_allclose(O.recursive_sta_lta(g["signal"], nsta, nlta), g["recursive"], rtol=1e-13)


def test_port_matches_golden_iceland(iceland):
    g = iceland
    tau = g["tau"].astype(np.int32)
    mx, nm, ix = O.detect_fused(g["onsets"], tau, int(g["fsmp"]), int(g["lsmp"]), int(g["available"]), threads=8)
    assert np.array_equal(ix, g["max_coa_idx"])
    np.testing.assert_allclose(mx, g["max_coa"], rtol=1e-14)
    np.testing.assert_allclose(nm, g["max_norm_coa"], rtol=1e-13)
    assert ix[300] == np.ravel_multi_index(tuple(g["planted_node"]), tau.shape[:-1])


# ---- GPU: CUDA path vs reference outputs ----------------------------------------------------------
@pytest.fixture(scope="module")
def engine():
    from quakemigrate_b200.core import Engine
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.mark.gpu
def test_cuda_matches_golden_small(engine, small):
    g = small
    fsmp, lsmp, S = int(g["fsmp"]), int(g["lsmp"]), int(g["available"])
    lut = engine.upload_lut(g["tau"])
    mx, nm, ix = engine.detect(lut, g["onsets"], fsmp, lsmp, S)
    assert np.array_equal(ix, g["max_coa_idx"])
    np.testing.assert_allclose(mx, g["max_coa"], rtol=1e-9)
    np.testing.assert_allclose(nm, g["max_norm_coa"], rtol=1e-5)
    dev = engine.upload_onsets(g["onsets"])
    T = g["onsets"].shape[1] - fsmp - lsmp
    coa, slab = engine.locate_window(lut, dev, fsmp, lsmp, S, 0, T, return_map4d=True)
    np.testing.assert_allclose(slab, g["map4d"], rtol=1e-5)
    np.testing.assert_allclose([slab.sum(), (slab ** 2).sum()], g["map4d_checksum"], rtol=1e-6)
    sig = g["signal"]
    nsta, nlta = int(g["nsta"]), int(g["nlta"])
    for pos, key in (("classic", "classic"), ("centred", "centred"), ("recursive", "recursive")):
        np.testing.assert_allclose(engine.sta_lta(sig, nsta, nlta, pos)[0], g[key], rtol=1e-9)
    dev.free()
    lut.free()


@pytest.mark.gpu
def test_cuda_matches_golden_iceland(engine, iceland):
    """BASELINE config 1 shape: real example LUT (decimated), 24 rows, 250 Hz, T=625."""
    g = iceland
    tau = g["tau"].astype(np.int32)
    lut = engine.upload_lut(tau)
    mx, nm, ix = engine.detect(lut, g["onsets"], int(g["fsmp"]), int(g["lsmp"]), int(g["available"]))
    assert np.array_equal(ix, g["max_coa_idx"])
    np.testing.assert_allclose(mx, g["max_coa"], rtol=1e-9)
    np.testing.assert_allclose(nm, g["max_norm_coa"], rtol=1e-5)
    # same through the reference-compatible T1 symbols
    from quakemigrate_b200 import core
    m4 = core.migrate(g["onsets"], tau, int(g["fsmp"]), int(g["lsmp"]), int(g["available"]), 1)
    f = core.find_max_coa(m4, 1)
    assert np.array_equal(f[2], g["max_coa_idx"])
    np.testing.assert_allclose(f[0], g["max_coa"], rtol=1e-5)
    lut.free()
