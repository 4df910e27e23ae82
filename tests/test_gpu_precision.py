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


@pytest.mark.parametrize("precision", ["q16", "fp32", "auto"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_all_operand_choices_meet_the_contract(precision, seed):
    """Unstructured stress inputs (random travel times, gamma onsets, a fifth of the samples on the
    clip floor, some negative travel times).  Indices and max_coa are exact whatever the operands;
    max_norm_coa is inside 1e-5 with float32 rows AND in the automatic mode (which verifies the
    fixed-point node sums and falls back when needed); with fixed point FORCED on these tiny grids
    its error is whatever the library's own check reports."""
    rng = np.random.default_rng(seed)
    grid = (int(rng.integers(9, 22)), int(rng.integers(9, 22)), int(rng.integers(5, 14)))
    S, T = int(rng.integers(3, 40)), int(rng.integers(2100, 3000))
    fsmp, lsmp = int(rng.integers(0, 50)), int(rng.integers(60, 200))
    onsets, tau = syn.random_problem(seed + 50, grid, S, T, fsmp, lsmp, neg_frac=0.03, floor_frac=0.2)
    want = O.detect_fused(onsets, tau, fsmp, lsmp, S, threads=4)
    got, info = _run_info(precision, onsets, tau, fsmp, lsmp, S)
    assert info["operand_bytes"] == {"q16": 2, "fp32": 4}.get(precision, info["operand_bytes"])
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    err = float(np.max(np.abs(got[1] - want[1]) / want[1]))
    if precision == "q16":
        assert err <= max(1e-5, 4.0 * info["check_max_rel_err"])
    else:
        assert err <= 1e-5
        assert info["operand_bytes"] == 4 or info["check_max_rel_err"] <= 2e-6


def _run_info(precision, onsets, tau, fsmp, lsmp, avail, qgroup=None):
    from quakemigrate_b200.core import Engine
    with Engine(0, precision=precision, qgroup=qgroup) as eng:
        lut = eng.upload_lut(tau)
        got = eng.detect(lut, onsets, fsmp, lsmp, avail)
        info = eng.last_scan_info
        lut.free()
    return got, info


def test_automatic_choice():
    """auto = float32 rows for short scans (the verification windows would cost more than the 16-bit
    rows save), 16-bit rows otherwise -- and then every call VERIFIES the node sums of the 16-bit
    scan against float32 re-scans of 1-2 windows and redoes the scan in float32 when they differ by
    more than 2e-6: max_norm_coa is inside 1e-5 whatever path was taken."""
    small = syn.random_problem(1, (3, 2, 2), 24, 300, 10, 60)
    got, info = _run_info("auto", small[0], small[1], 10, 60, 24)
    assert info["operand_bytes"] == 4 and info["check_windows"] == 0
    want = O.detect_fused(small[0], small[1], 10, 60, 24, threads=2)
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5)

    cfg = syn.CONFIGS["cfg2"].scaled(grid=(40, 36, 20), n_samples=2400, n_rows=16, name="auto")
    prob = syn.make_problem(cfg)
    big = (syn.onsets_from_waveforms(prob, lambda x, a, b: O.overlapping_sta_lta(x, a, b)), prob.traveltimes)
    fsmp, lsmp = prob.fsmp, prob.lsmp
    got, info = _run_info("auto", big[0], big[1], fsmp, lsmp, 16)
    assert info["operand_bytes"] == 2 and info["fell_back_fp32"] == 0 and info["operand_bits"] in (13, 14, 15, 16)
    assert info["check_windows"] >= 1 and info["check_samples"] >= 200
    assert info["check_max_rel_err"] <= 2e-6
    assert info["host_syncs_in_timed_region"] == 0
    want = O.detect_fused(big[0], big[1], fsmp, lsmp, 16, threads=8)
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5)

    # available = 1: the quantisation error of a stack is not divided by the row count any more;
    # the check sees it and the result comes from float32 rows
    got, info = _run_info("auto", big[0], big[1], fsmp, lsmp, 1)
    want = O.detect_fused(big[0], big[1], fsmp, lsmp, 1, threads=8)
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5)
    assert info["operand_bytes"] == 4 or info["check_max_rel_err"] <= 2e-6


@pytest.mark.parametrize("qgroup,bits", [(1, 16), (2, 15), (4, 14), (8, 13)])
def test_every_row_group_meets_the_contract(qgroup, bits):
    """The packed kernels (2, 4, 8 rows per carry-free group on 15-, 14-, 13-bit rows) and the
    two-instruction 16-bit kernel give the same answers: indices exact, max_coa 1e-9, the node sum
    verified.  Row counts that are not multiples of the group exercise the zero pad rows."""
    for S, grid, T in ((13, (22, 19, 11), 2300), (32, (30, 26, 12), 2100), (5, (17, 12, 9), 2600)):
        cfg = syn.CONFIGS["cfg2"].scaled(grid=grid, n_samples=T, n_rows=S, name="grp")
        prob = syn.make_problem(cfg)
        onsets = syn.onsets_from_waveforms(prob, lambda x, a, b: O.overlapping_sta_lta(x, a, b))
        got, info = _run_info("q16", onsets, prob.traveltimes, prob.fsmp, prob.lsmp, S, qgroup=qgroup)
        assert info["operand_bits"] == bits and info["operand_bytes"] == 2
        want = O.detect_fused(onsets, prob.traveltimes, prob.fsmp, prob.lsmp, S, threads=8)
        assert np.array_equal(got[2], want[2]), (qgroup, S)
        np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
        np.testing.assert_allclose(got[1], want[1], rtol=1e-4)  # forced fixed point on tiny grids
        assert abs(info["check_max_rel_err"] - np.max(np.abs(got[1] - want[1]) / want[1])) < 1e-4


def _adversarial_rows(S, P, kind, rng, bits=16):
    """Linear onsets whose LOG values sit half a quantisation step off the 16-bit grid, so that the
    rounding residuals of all rows and all nodes have the same sign (the case a sqrt(N) averaging
    argument gets wrong).  Two samples per row pin the common range [log 0.4, log 50]."""
    lo, hi = np.log(0.4), np.log(50.0)
    step = (hi - lo) / float(2 ** bits - 1)
    sc = 2 ** (16 - bits)
    t = np.arange(P)
    rows = np.empty((S, P))
    for s in range(S):
        if kind == "constant":
            level = np.full(P, (20000 + 37 * s) // sc)
        elif kind == "ramp":          # constant plus a slow ramp: a few thousand distinct levels
            level = (12000 + 11 * s) // sc + t // 3
        else:                          # "steps": few distinct values
            level = (15000 + 1000 * ((t // 50 + s) % 4)) // sc
        rows[s] = np.exp(lo + (level + 0.5 - 1e-3) * step)
        rows[s, rng.integers(0, P)] = 0.4
        rows[s, rng.integers(0, P)] = 50.0
    return rows


def _assert_argmax_equal_up_to_rounding_ties(idx, map4d, rel=1e-13):
    """These rows take a few thousand distinct values, so different nodes reach the SAME stack up to
    the last bits of the float64 sum; which of them has the larger exp() then depends on the libm
    (the reference itself is built -Ofast).  Where the indices differ, the node we report must hold
    the reference's maximum to `rel`."""
    flat = map4d.reshape(-1, map4d.shape[-1])
    want = np.argmax(flat, axis=0)
    bad = np.nonzero(idx != want)[0]
    if len(bad):
        ours = flat[idx[bad], bad]
        assert np.all(ours >= flat[want[bad], bad] * (1.0 - rel)), "arg-max differs beyond rounding-level ties"
    return len(bad)


@pytest.mark.parametrize("kind", ["constant", "ramp", "steps"])
@pytest.mark.parametrize("few_tau", [False, True])
@pytest.mark.parametrize("qgroup,bits", [(1, 16), (2, 15)])
def test_correlated_quantisation_residuals_are_caught(kind, few_tau, qgroup, bits):
    """Adversarial inputs for the 16-bit node sum: every residual is ~ +step/2,
    so the quantised node sum is off by S * step / (2 available) ~ 3.7e-5 relative (16-bit rows; twice
    that per bit dropped) -- outside the bar.  The automatic mode must notice (check_max_rel_err) and deliver 1e-5 anyway; arg-max and
    max_coa stay exact in either format."""
    rng = np.random.default_rng(7)
    grid, S, T, fsmp, lsmp = (20, 18, 10), 8, 2500, 10, 80
    P = fsmp + T + lsmp
    onsets = _adversarial_rows(S, P, kind, rng, bits)
    if few_tau:
        tau = rng.integers(0, 3, size=grid + (S,), dtype=np.int32) * 20
    else:
        _, tau = syn.random_problem(5, grid, S, T, fsmp, lsmp)
    map4d = O.migrate(onsets, tau, fsmp, lsmp, S, threads=8)
    want = O.find_max_coa(map4d, threads=8)
    got, info = _run_info("auto", onsets, tau, fsmp, lsmp, S, qgroup=qgroup)
    _assert_argmax_equal_up_to_rounding_ties(got[2], map4d)
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5)
    assert info["check_windows"] >= 1
    assert info["fell_back_fp32"] == 1 and info["operand_bytes"] == 4 and info["check_max_rel_err"] > 2e-6
    # forced 16-bit rows: same indices and max_coa, and the reported check error IS the error
    got_q, info_q = _run_info("q16", onsets, tau, fsmp, lsmp, S, qgroup=qgroup)
    _assert_argmax_equal_up_to_rounding_ties(got_q[2], map4d)
    np.testing.assert_allclose(got_q[0], want[0], rtol=1e-9)
    err = float(np.max(np.abs(got_q[1] - want[1]) / want[1]))
    assert info_q["operand_bytes"] == 2 and info_q["fell_back_fp32"] == 0 and info_q["operand_bits"] == bits
    assert err > 1e-5 and 0.3 * err <= info_q["check_max_rel_err"] <= 3.0 * err


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
    np.testing.assert_allclose(res["q16"][1], res["fp32"][1], rtol=5e-5)


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
    np.testing.assert_allclose(got[1], want[1], rtol=4e-3)  # step = log(5e6 / 0.01) / 32767 on 960 nodes
    # the automatic mode sees that and delivers 1e-5 from float32 rows (T >= 2048 to get there)
    onsets2, tau2 = syn.random_problem(9, grid, S, 2200, fsmp, lsmp)
    onsets2[3, 200:230] = 5.0e6
    want2 = O.detect_fused(onsets2, tau2, fsmp, lsmp, S, threads=4)
    got2, info2 = _run_info("auto", onsets2, tau2, fsmp, lsmp, S)
    assert np.array_equal(got2[2], want2[2])
    np.testing.assert_allclose(got2[1], want2[1], rtol=1e-5)
    assert info2["operand_bytes"] == 4 or info2["check_max_rel_err"] <= 2e-6
