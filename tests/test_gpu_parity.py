# -*- coding: utf-8 -*-
"""
GPU parity tests: libqmb200 (through its C ABI / ctypes) against the CPU oracle on identical
seeded inputs.  Bars (north_star): arg-max indices bit-exact; coalescence values within 1e-5
relative (we assert much tighter where the arithmetic allows).
"""

import numpy as np
import pytest

from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

RTOL_COA = 1e-5          # stated tolerance of the north star
RTOL_COA_TIGHT = 1e-9    # max_coa is re-evaluated in float64: expect ~1e-15
RTOL_NORM = 1e-5         # north_star bar; the node sum is built from 16-bit fixed-point (or fp32) stacks: 3e-6 observed on tiny grids, ~1e-7 at bench sizes


def _oracle_kind():
    return "ref" if O.have("ref") else "port"


@pytest.fixture(scope="module")
def engine():
    from quakemigrate_b200.core import Engine
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="module", params=list(range(10)))
def engine_v(request):
    """Every kernel tiling variant (QMB_VARIANTS in csrc/qmb_internal.cuh)."""
    from quakemigrate_b200.core import Engine
    eng = Engine(0, variant=request.param)
    yield eng
    eng.close()


def _check_detect(got, want, rtol_norm=RTOL_NORM):
    g_max, g_norm, g_idx = got
    w_max, w_norm, w_idx = want
    assert np.array_equal(g_idx, w_idx), f"argmax mismatch at {np.nonzero(g_idx != w_idx)[0][:10]}"
    np.testing.assert_allclose(g_max, w_max, rtol=RTOL_COA_TIGHT)
    np.testing.assert_allclose(g_norm, w_norm, rtol=rtol_norm)


# ------------------------------------------------------------------------------------------------
# STA/LTA
# ------------------------------------------------------------------------------------------------
def test_stalta_reference_kat():
    """tests/test_onsets.py:23-39 of the reference, through the T1 symbols."""
    from quakemigrate_b200 import core
    signal = np.arange(6)
    overlapping = core.overlapping_sta_lta(signal, 2, 3)
    assert (overlapping == np.array([1.0, 1.0, 1.5, 1.25, 21. / 18, 27. / 24])).all()
    centred = core.centred_sta_lta(signal, 2, 3)
    assert np.allclose(centred, np.array([1.0, 1.0, 3.5, 2.25, 1.0, 1.0]))


@pytest.mark.parametrize("position", ["classic", "centred", "recursive"])
@pytest.mark.parametrize("n,nsta,nlta", [(5000, 21, 101), (4097, 3, 251), (310, 5, 300), (64, 1, 2)])
def test_stalta_random(engine, position, n, nsta, nlta):
    rng = np.random.default_rng(n + nsta)
    sig = rng.standard_normal((3, n)) ** 2
    got = engine.sta_lta(sig, nsta, nlta, position)
    fn = {"classic": O.overlapping_sta_lta, "centred": O.centred_sta_lta,
          "recursive": O.recursive_sta_lta}[position]
    want = np.array([fn(s, nsta, nlta, kind=_oracle_kind()) for s in sig])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)


def test_stalta_centred_zero_lta(engine):
    """lta <= 0 -> exactly 1.0 for i >= nlta (onsetlib.c:102-106)."""
    sig = np.zeros((1, 400))
    sig[0, 200:] = 1.0
    got = engine.sta_lta(sig, 5, 50, "centred")
    want = O.centred_sta_lta(sig[0], 5, 50, kind=_oracle_kind())
    sel = np.isfinite(want)
    np.testing.assert_allclose(got[0][sel], want[sel], rtol=1e-12)
    assert np.array_equal(np.isfinite(got[0]), sel)


def test_onsets_from_waveforms(engine):
    cfg = syn.CONFIGS["cfg2"].scaled(grid=(6, 6, 4), n_samples=3000, n_rows=6)
    prob = syn.make_problem(cfg)
    dev, lin = engine.onsets_from_waveforms(prob.waveforms, prob.row_channels, prob.nsta, prob.nlta,
                                            position="classic", return_onsets=True)
    want = syn.onsets_from_waveforms(prob, lambda x, a, b: O.overlapping_sta_lta(x, a, b, kind=_oracle_kind()))
    np.testing.assert_allclose(lin, want, rtol=1e-9)
    dev.free()


# ------------------------------------------------------------------------------------------------
# detect: fused migrate + find_max_coa
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid,S,T,fsmp,lsmp", [
    ((6, 5, 4), 7, 50, 13, 40),        # tiny, ragged tiles, odd rows
    ((9, 8, 10), 16, 300, 0, 64),      # fsmp = 0
    ((13, 4, 4), 33, 513, 7, 31),      # three row chunks, T not multiple of the tile
    ((4, 4, 4), 1, 1, 1, 5),           # single row, single sample
    ((5, 5, 5), 128, 300, 4, 40),      # BASELINE cfg5 row count (128 rows)
    ((6, 4, 7), 64, 257, 0, 33),       # BASELINE cfg3 row count
    ((3, 2, 2), 24, 640, 413, 1000),   # BASELINE cfg1 pads (fsmp 413, lsmp 1000), grid < one tile
])
def test_detect_random(engine_v, grid, S, T, fsmp, lsmp):
    onsets, tau = syn.random_problem(S * 1000 + T, grid, S, T, fsmp, lsmp, neg_frac=0.05)
    lut = engine_v.upload_lut(tau)
    got = engine_v.detect(lut, onsets, fsmp, lsmp, S)
    m4 = O.migrate(onsets, tau, fsmp, lsmp, S, threads=4, kind=_oracle_kind())
    want = O.find_max_coa(m4, threads=4, kind=_oracle_kind())
    _check_detect(got, want)
    lut.free()


def test_detect_synthetic_homogeneous(engine_v):
    cfg = syn.CONFIGS["cfg2"].scaled(grid=(21, 18, 11), n_samples=1500, n_rows=8)
    prob = syn.make_problem(cfg)
    onsets = syn.onsets_from_waveforms(prob, lambda x, a, b: O.overlapping_sta_lta(x, a, b))
    lut = engine_v.upload_lut(prob.traveltimes)
    got = engine_v.detect(lut, onsets, prob.fsmp, prob.lsmp, cfg.n_rows)
    want = O.detect_fused(onsets, prob.traveltimes, prob.fsmp, prob.lsmp, cfg.n_rows, threads=8)
    _check_detect(got, want)
    # the planted events must show up as maxima near their origin time at the right node
    lut.free()


def test_detect_exact_ties_resolve_to_lowest_index(engine_v):
    """Constant onsets: every node ties at every sample -> index 0 (migratelib.c:96-105)."""
    grid, S, T, fsmp, lsmp = (5, 6, 7), 5, 70, 3, 20
    onsets = np.full((S, fsmp + T + lsmp), 0.4)
    tau = np.random.default_rng(1).integers(0, lsmp + 1, size=grid + (S,), dtype=np.int32)
    lut = engine_v.upload_lut(tau)
    g_max, g_norm, g_idx = engine_v.detect(lut, onsets, fsmp, lsmp, S)
    assert (g_idx == 0).all()
    np.testing.assert_allclose(g_max, 0.4, rtol=1e-12)
    np.testing.assert_allclose(g_norm, 1.0, rtol=1e-6)
    lut.free()


def test_detect_partial_ties(engine_v):
    """Many values sit exactly on the clip floor: pair/slow-path candidates must match the oracle."""
    grid, S, T, fsmp, lsmp = (8, 8, 8), 4, 200, 5, 30
    onsets, tau = syn.random_problem(77, grid, S, T, fsmp, lsmp, floor_frac=0.9)
    lut = engine_v.upload_lut(tau)
    got = engine_v.detect(lut, onsets, fsmp, lsmp, S)
    want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S, threads=4), threads=4)
    _check_detect(got, want)
    lut.free()


def test_detect_errors(engine):
    onsets, tau = syn.random_problem(3, (4, 4, 4), 3, 20, 2, 10)
    lut = engine.upload_lut(tau)
    with pytest.raises(ValueError):   # lib.py:105-108
        engine.detect(lut, onsets[:2], 2, 10, 3)
    with pytest.raises(ValueError):   # lib.py:109-110
        engine.detect(lut, onsets, 20, 20, 3)
    from quakemigrate_b200.core import QmbError
    with pytest.raises(QmbError):     # tau beyond the post-pad: the reference would read past rows
        engine.detect(lut, onsets[:, :-5], 2, 5, 3)
    lut.free()


# ------------------------------------------------------------------------------------------------
# T1: reference-compatible symbols
# ------------------------------------------------------------------------------------------------
def test_legacy_migrate_find_max(engine):
    from quakemigrate_b200 import core
    grid, S, T, fsmp, lsmp = (7, 6, 5), 6, 90, 4, 25
    onsets, tau = syn.random_problem(11, grid, S, T, fsmp, lsmp, neg_frac=0.1)
    m_got = core.migrate(onsets, tau, fsmp, lsmp, S, 1)
    m_want = O.migrate(onsets, tau, fsmp, lsmp, S, threads=2, kind=_oracle_kind())
    assert m_got.shape == m_want.shape
    np.testing.assert_allclose(m_got, m_want, rtol=RTOL_COA)
    f_got = core.find_max_coa(m_want, 1)
    f_want = O.find_max_coa(m_want, threads=2, kind=_oracle_kind())
    assert np.array_equal(f_got[2], f_want[2])
    np.testing.assert_allclose(f_got[0], f_want[0], rtol=0, atol=0)
    np.testing.assert_allclose(f_got[1], f_want[1], rtol=1e-13)
    with pytest.raises(ValueError):
        core.migrate(onsets[:3], tau, fsmp, lsmp, S, 1)


# ------------------------------------------------------------------------------------------------
# locate: marginal window
# ------------------------------------------------------------------------------------------------
def test_locate_window(engine_v):
    grid, S, T, fsmp, lsmp = (9, 7, 6), 6, 400, 10, 40
    onsets, tau = syn.random_problem(5, grid, S, T, fsmp, lsmp)
    lut = engine_v.upload_lut(tau)
    dev = engine_v.upload_onsets(onsets)
    m4 = O.migrate(onsets, tau, fsmp, lsmp, S, threads=4, kind=_oracle_kind())
    for w0, w1 in [(0, T), (123, 301), (399, 400)]:
        coa, slab = engine_v.locate_window(lut, dev, fsmp, lsmp, S, w0, w1, return_map4d=True)
        np.testing.assert_allclose(slab, m4[..., w0:w1], rtol=RTOL_COA)
        np.testing.assert_allclose(coa, m4[..., w0:w1].sum(axis=-1), rtol=RTOL_COA)
    dev.free()
    lut.free()


def test_detect_random_shapes(engine):
    """Seeded sweep over awkward shapes: tiny grids, T < one lane group, S = 1, zero pads."""
    rng = np.random.default_rng(20260925)
    for case in range(40):
        grid = tuple(int(x) for x in rng.integers(1, 10, size=3))
        S = int(rng.integers(1, 41))
        T = int(rng.choice([1, 2, 31, 32, 33, 255, 256, 257, int(rng.integers(1, 700))]))
        fsmp = int(rng.choice([0, 1, 2, 3, int(rng.integers(0, 50))]))
        lsmp = int(rng.integers(0, 80))
        onsets, tau = syn.random_problem(1000 + case, grid, S, T, fsmp, lsmp,
                                         neg_frac=float(rng.choice([0.0, 0.1])),
                                         floor_frac=float(rng.choice([0.0, 0.3, 0.95])))
        lut = engine.upload_lut(tau)
        got = engine.detect(lut, onsets, fsmp, lsmp, S)
        want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S, threads=2), threads=2)
        assert np.array_equal(got[2], want[2]), (case, grid, S, T, fsmp, lsmp)
        np.testing.assert_allclose(got[0], want[0], rtol=RTOL_COA_TIGHT, err_msg=str((case, grid, S, T)))
        np.testing.assert_allclose(got[1], want[1], rtol=RTOL_NORM, err_msg=str((case, grid, S, T)))
        lut.free()


def test_available_differs_from_row_count(engine):
    """exp(stack / available) uses `available`, not the row count (migratelib.c:62)."""
    grid, S, T, fsmp, lsmp = (5, 4, 6), 6, 120, 3, 25
    onsets, tau = syn.random_problem(5, grid, S, T, fsmp, lsmp)
    lut = engine.upload_lut(tau)
    for avail in (1, 4, 9):
        got = engine.detect(lut, onsets, fsmp, lsmp, avail)
        want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, avail, threads=2), threads=2)
        _check_detect(got, want)
    lut.free()


def test_fine_grid_with_many_shared_offsets(engine):
    """Fine grid / low sampling rate: most neighbouring nodes share travel-time offsets."""
    cfg = syn.CONFIGS["cfg1fine"].scaled(grid=(23, 20, 19), n_samples=700, n_rows=8)
    prob = syn.make_problem(cfg)
    onsets = syn.onsets_from_waveforms(prob, lambda x, a, b: O.overlapping_sta_lta(x, a, b))
    lut = engine.upload_lut(prob.traveltimes)
    assert lut.info()["duplicate_fraction"] > 0.3
    got = engine.detect(lut, onsets, prob.fsmp, prob.lsmp, cfg.n_rows)
    want = O.detect_fused(onsets, prob.traveltimes, prob.fsmp, prob.lsmp, cfg.n_rows, threads=4)
    _check_detect(got, want)
    lut.free()


def test_onset_clip_floor_and_log(engine):
    """np.clip(onsets, 0.01, inf) before the log (lib.py:93-94): zeros, negatives and tiny values."""
    grid, S, T, fsmp, lsmp = (5, 5, 4), 4, 90, 2, 20
    onsets, tau = syn.random_problem(8, grid, S, T, fsmp, lsmp)
    rng = np.random.default_rng(8)
    sel = rng.random(onsets.shape)
    onsets[sel < 0.10] = 0.0
    onsets[(sel >= 0.10) & (sel < 0.15)] = -3.0
    onsets[(sel >= 0.15) & (sel < 0.20)] = 1e-7
    onsets[(sel >= 0.20) & (sel < 0.25)] = 0.01
    lut = engine.upload_lut(tau)
    got = engine.detect(lut, onsets, fsmp, lsmp, S)
    want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S, threads=2), threads=2)
    _check_detect(got, want)
    lut.free()


@pytest.mark.parametrize("n_levels,jitter", [(3, 1e-9), (2, 3e-8), (6, 1e-12)])
def test_near_ties_below_fp32_resolution(engine, n_levels, jitter):
    """
    Onsets drawn from a few discrete levels times (1 + k*jitter): many nodes have stacks that
    differ only far below float32 resolution (but not exactly), so the fp32 pass cannot rank them
    and the float64 pair / slow paths must reproduce the reference's arg-max exactly.
    """
    grid, S, T, fsmp, lsmp = (7, 6, 6), 5, 180, 4, 24
    rng = np.random.default_rng(n_levels * 1000 + int(-np.log10(jitter)))
    levels = np.array([0.5, 0.8, 1.3, 2.0, 3.1, 4.7])[:n_levels]
    total = fsmp + T + lsmp
    onsets = levels[rng.integers(0, n_levels, size=(S, total))] * (1.0 + jitter * rng.integers(0, 4, size=(S, total)))
    tau = rng.integers(0, lsmp + 1, size=grid + (S,), dtype=np.int32)
    lut = engine.upload_lut(tau)
    info_ms = engine.detect_device(lut, engine.upload_onsets(onsets), fsmp, lsmp, S)
    got = engine.detect_fetch(T)
    want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S, threads=2), threads=2)
    assert info_ms["slow_samples"] > 0          # the case really exercises the slow path
    _check_detect(got, want)
    lut.free()


def test_slow_path_partial_maximum_filter(engine, monkeypatch):
    """
    The all-node float64 re-scan only walks node subsets whose partial maximum of the scan reaches
    m1 - thr.  Same input with the filter off (QMB_SLOW_NOFILTER=1): identical results, both equal
    to the oracle; several node ranges and tiles so that the filter really discards subsets.
    """
    grid, S, T, fsmp, lsmp = (24, 20, 18), 6, 700, 4, 30
    rng = np.random.default_rng(77)
    levels = np.array([0.5, 0.8, 1.3])
    total = fsmp + T + lsmp
    onsets = levels[rng.integers(0, 3, size=(S, total))] * (1.0 + 1e-9 * rng.integers(0, 4, size=(S, total)))
    ix, iy, iz = np.meshgrid(*[np.arange(g) for g in grid], indexing="ij")
    tau = np.stack([np.round(0.7 * np.sqrt((ix - 3 * s) ** 2 + (iy - 2 * s) ** 2 + (iz + s) ** 2))
                    for s in range(S)], axis=-1).astype(np.int32)
    assert tau.max() <= lsmp
    lut = engine.upload_lut(tau)
    dev_on = engine.upload_onsets(onsets)
    info = engine.detect_device(lut, dev_on, fsmp, lsmp, S)
    got = engine.detect_fetch(T)
    assert info["slow_samples"] > 0
    monkeypatch.setenv("QMB_SLOW_NOFILTER", "1")
    info2 = engine.detect_device(lut, dev_on, fsmp, lsmp, S)
    got2 = engine.detect_fetch(T)
    assert info2["slow_samples"] == info["slow_samples"]
    for a, b in zip(got, got2):
        np.testing.assert_array_equal(a, b)
    want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S, threads=2), threads=2)
    _check_detect(got, want)
    lut.free()


def test_span_too_wide_is_an_error_not_a_fallback(engine):
    """Travel-time range inside one tile beyond what shared memory can stage -> QMB_ERR_UNSUPPORTED."""
    from quakemigrate_b200.core import QmbError
    grid, S = (4, 4, 4), 3
    tau = np.zeros(grid + (S,), dtype=np.int32)
    tau[0, 0, 0, :] = 20000                       # one node 20000 samples away from its neighbours
    with pytest.raises(QmbError, match="too wide|too large"):
        engine.upload_lut(tau)


def test_long_trace_small_grid(engine):
    """Hundreds of thousands of samples: index arithmetic, many time tiles, few node tiles."""
    grid, S, T, fsmp, lsmp = (3, 3, 3), 3, 300_001, 7, 40
    onsets, tau = syn.random_problem(12, grid, S, T, fsmp, lsmp)
    lut = engine.upload_lut(tau)
    got = engine.detect(lut, onsets, fsmp, lsmp, S)
    want = O.detect_fused(onsets, tau, fsmp, lsmp, S, threads=O.default_threads())
    _check_detect(got, want)
    lut.free()
