# -*- coding: utf-8 -*-
"""
The detect scan gathers 16-bit fixed-point onset rows by default (two samples per shared-memory
word, exact integer stacks of the quantised rows) and float32 rows on request.  Whatever the
operand format, the winner of every sample is decided in float64 among the candidates inside a
rigorous error band, so the contract is the same: arg-max indices bit-exact against the
reference's migrate + find_max_coa (migratelib.c:40-111), max_coa to 1e-9, max_norm_coa within
the 1e-5 relative bar of BASELINE.json's north_star.
"""
import numpy as np
import pytest

from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _run(precision, onsets, tau, fsmp, lsmp, avail):
    from quakemigrate_b200.core import Engine
    with Engine(0, precision=precision) as eng:
        lut = eng.upload_lut(tau)
        got = eng.detect(lut, onsets, fsmp, lsmp, avail)
        bytes_per_gather = eng.last_scan_bytes_per_gather
        lut.free()
    return got, bytes_per_gather


@pytest.mark.parametrize("precision,bpg", [("q16", 2), ("fp32", 4)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_both_operand_formats_meet_the_contract(precision, bpg, seed):
    rng = np.random.default_rng(seed)
    grid = (int(rng.integers(9, 22)), int(rng.integers(9, 22)), int(rng.integers(5, 14)))
    S, T = int(rng.integers(3, 40)), int(rng.integers(300, 1500))
    fsmp, lsmp = int(rng.integers(0, 50)), int(rng.integers(60, 200))
    onsets, tau = syn.random_problem(seed + 50, grid, S, T, fsmp, lsmp, neg_frac=0.03, floor_frac=0.2)
    want = O.detect_fused(onsets, tau, fsmp, lsmp, S, threads=4)
    got, used = _run(precision, onsets, tau, fsmp, lsmp, S)
    assert used == bpg
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    # forced 16-bit rows on a small grid: the node sum averages the quantisation noise over few
    # nodes only (the automatic mode keeps float32 here, see test_automatic_choice)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5 if precision == "fp32" else 1e-4)


def test_automatic_choice():
    """auto = 16-bit rows at realistic sizes, float32 rows where the node sum could not average the
    quantisation noise down (tiny grids; `available` far below the row count)."""
    small = syn.random_problem(1, (3, 2, 2), 24, 300, 10, 60)
    got, used = _run("auto", small[0], small[1], 10, 60, 24)
    assert used == 4
    want = O.detect_fused(small[0], small[1], 10, 60, 24, threads=2)
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5)
    big = syn.random_problem(2, (40, 40, 30), 16, 400, 10, 60)
    got, used = _run("auto", big[0], big[1], 10, 60, 16)
    assert used == 2
    want = O.detect_fused(big[0], big[1], 10, 60, 16, threads=8)
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5)
    got, used = _run("auto", big[0], big[1], 10, 60, 1)  # available = 1: no division of the error
    assert used == 4


def test_formats_agree_with_each_other_on_a_planted_event():
    cfg = syn.CONFIGS["cfg2"].scaled(grid=(24, 20, 12), n_samples=2000, n_rows=16, name="prec")
    prob = syn.make_problem(cfg)
    from quakemigrate_b200.core import Engine
    res = {}
    for precision in ("q16", "fp32"):
        with Engine(0, precision=precision) as eng:
            dev = eng.onsets_from_waveforms(prob.waveforms, prob.row_channels, prob.nsta, prob.nlta,
                                            position=cfg.position)
            lut = eng.upload_lut(prob.traveltimes)
            eng.detect_device(lut, dev, prob.fsmp, prob.lsmp, cfg.n_rows)
            res[precision] = eng.detect_fetch(cfg.n_samples)
    assert np.array_equal(res["q16"][2], res["fp32"][2])
    np.testing.assert_allclose(res["q16"][0], res["fp32"][0], rtol=1e-12)
    np.testing.assert_allclose(res["q16"][1], res["fp32"][1], rtol=2e-5)


def test_clip_floor_is_on_the_quantisation_grid():
    """Rows made of the clip floor only (onset = 0.4 everywhere): every stack is exact, all nodes
    tie, the first node wins (migratelib.c:96-105) and max_norm_coa is exactly 1."""
    grid, S, T, fsmp, lsmp = (8, 8, 6), 7, 300, 10, 40
    _, tau = syn.random_problem(3, grid, S, T, fsmp, lsmp)
    onsets = np.full((S, fsmp + T + lsmp), 0.4)
    got, used = _run("q16", onsets, tau, fsmp, lsmp, S)
    assert used == 2
    assert (got[2] == 0).all()
    np.testing.assert_allclose(got[0], 0.4, rtol=1e-12)
    np.testing.assert_allclose(got[1], 1.0, rtol=1e-6)  # exp2 of the float32 stack, summed over nodes


def test_wide_dynamic_range_rows():
    """One huge transient stretches the quantisation step; indices must stay exact."""
    grid, S, T, fsmp, lsmp = (12, 10, 8), 12, 600, 20, 80
    onsets, tau = syn.random_problem(9, grid, S, T, fsmp, lsmp)
    onsets[3, 200:230] = 5.0e6
    onsets[7, 400:410] = 1.0e-9  # below the 0.01 clip of lib.py:93
    want = O.detect_fused(onsets, tau, fsmp, lsmp, S, threads=4)
    got, _ = _run("q16", onsets, tau, fsmp, lsmp, S)
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-3)  # step = log(5e6 / 0.01) / 65535 on 960 nodes
