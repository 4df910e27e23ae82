# -*- coding: utf-8 -*-
"""
CPU tests pinning the oracle (our restatement) against the reference: its own KAT vectors, the
reference C compiled here (oracle/_ref, when present) and the committed golden vectors.
"""

import numpy as np
import pytest

from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn

KINDS = ["port"] + [k for k in ("ref", "ref_strict") if O.have(k)]


@pytest.mark.parametrize("kind", KINDS)
def test_reference_kat(kind):
    """tests/test_onsets.py:23-39."""
    signal = np.arange(6)
    overlapping = O.overlapping_sta_lta(signal, 2, 3, kind=kind)
    assert (overlapping == np.array([1.0, 1.0, 1.5, 1.25, 21. / 18, 27. / 24])).all()
    centred = O.centred_sta_lta(signal, 2, 3, kind=kind)
    assert np.allclose(centred, np.array([1.0, 1.0, 3.5, 2.25, 1.0, 1.0]))


def test_numpy_restatement_kat():
    signal = np.arange(6)
    assert (O.np_overlapping_sta_lta(signal, 2, 3) == np.array([1.0, 1.0, 1.5, 1.25, 21. / 18, 27. / 24])).all()
    assert np.allclose(O.np_centred_sta_lta(signal, 2, 3), [1.0, 1.0, 3.5, 2.25, 1.0, 1.0])


@pytest.mark.skipif(not O.have("ref_strict"), reason="reference not compiled here")
def test_port_bit_identical_to_strict_reference():
    onsets, tau = syn.random_problem(0, (6, 5, 4), 7, 50, 13, 40, neg_frac=0.05)
    a = O.migrate(onsets, tau, 13, 40, 7, threads=3, kind="port")
    b = O.migrate(onsets, tau, 13, 40, 7, threads=3, kind="ref_strict")
    assert np.array_equal(a, b)
    for x, y in zip(O.find_max_coa(a, 2, "port"), O.find_max_coa(b, 2, "ref_strict")):
        assert np.array_equal(x, y)
    rng = np.random.default_rng(1)
    sig = rng.standard_normal(3000) ** 2
    for fn in ("overlapping_sta_lta", "centred_sta_lta", "recursive_sta_lta"):
        assert np.array_equal(getattr(O, fn)(sig, 21, 101, kind="port"),
                              getattr(O, fn)(sig, 21, 101, kind="ref_strict"))


@pytest.mark.skipif(not O.have("ref"), reason="reference not compiled here")
def test_port_vs_ofast_reference():
    """The reference's own -Ofast build differs from strict IEEE by a few ulp only."""
    onsets, tau = syn.random_problem(2, (8, 7, 6), 12, 120, 9, 50)
    a = O.migrate(onsets, tau, 9, 50, 12, threads=4, kind="port")
    b = O.migrate(onsets, tau, 9, 50, 12, threads=4, kind="ref")
    np.testing.assert_allclose(a, b, rtol=1e-14)
    fa, fb = O.find_max_coa(a, 4, "port"), O.find_max_coa(a, 4, "ref")
    assert np.array_equal(fa[2], fb[2])
    np.testing.assert_allclose(fa[0], fb[0], rtol=0)
    np.testing.assert_allclose(fa[1], fb[1], rtol=1e-13)


def test_numpy_restatement_matches_port():
    onsets, tau = syn.random_problem(4, (5, 4, 3), 5, 40, 6, 20, neg_frac=0.1)
    a = O.migrate(onsets, tau, 6, 20, 5, kind="port")
    b = O.np_migrate(onsets, tau, 6, 20, 5)
    np.testing.assert_allclose(a, b, rtol=1e-14)
    fa, fb = O.find_max_coa(a, 1, "port"), O.np_find_max_coa(a)
    assert np.array_equal(fa[2], fb[2])
    assert np.array_equal(fa[0], fb[0])
    np.testing.assert_allclose(fa[1], fb[1], rtol=1e-14)
    rng = np.random.default_rng(5)
    sig = rng.standard_normal(500) ** 2
    np.testing.assert_allclose(O.overlapping_sta_lta(sig, 7, 40), O.np_overlapping_sta_lta(sig, 7, 40), rtol=1e-10)
    np.testing.assert_allclose(O.centred_sta_lta(sig, 7, 40), O.np_centred_sta_lta(sig, 7, 40), rtol=1e-10)
    np.testing.assert_allclose(O.recursive_sta_lta(sig, 7, 40), O.np_recursive_sta_lta(sig, 7, 40), rtol=1e-12)


def test_fused_detect_is_bit_identical_to_two_step():
    onsets, tau = syn.random_problem(6, (7, 6, 5), 9, 300, 4, 33, neg_frac=0.02)
    two = O.find_max_coa(O.migrate(onsets, tau, 4, 33, 9, threads=2), threads=2)
    one = O.detect_fused(onsets, tau, 4, 33, 9, threads=3)
    for x, y in zip(one, two):
        assert np.array_equal(x, y)


def test_first_index_on_ties_and_accumulate_semantics():
    # all-equal onsets -> every node ties -> index 0 (migratelib.c:96-105)
    onsets = np.full((3, 40), 0.4)
    tau = np.random.default_rng(0).integers(0, 10, size=(3, 3, 3, 3), dtype=np.int32)
    m = O.migrate(onsets, tau, 5, 10, 3)
    mx, nm, ix = O.find_max_coa(m, 1)
    assert (ix == 0).all() and np.allclose(mx, 0.4) and np.allclose(nm, 1.0)
    # value errors of lib.py:105-110
    with pytest.raises(ValueError):
        O.migrate(onsets[:2], tau, 5, 10, 3)
    with pytest.raises(ValueError):
        O.migrate(onsets, tau, 30, 30, 3)
