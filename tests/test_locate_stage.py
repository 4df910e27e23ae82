# -*- coding: utf-8 -*-
"""
Locate, 3-D map stage: the oracle restatement and the CUDA kernels against
golden vectors produced by the reference's own _gaufilt3d / _gaufit3d / _covfit3d / _splineloc
(tests/golden/make_golden_locate.py lifts them out of scan.py with ``ast``).
"""

import pathlib

import numpy as np
import pytest

from oracle import qm_oracle as O

G = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "golden_locate.npz")
TAGS = ["a", "b"]


class _LUT:
    def __init__(self, shape, spacing, ll):
        self.node_count, self.node_spacing, self.ll_corner = np.array(shape), spacing, ll
        self.grid2coord = None

    def index2coord(self, value):
        return np.vstack(self.ll_corner + np.array(value) * self.node_spacing)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_restatement_matches_reference(tag):
    coa = G[f"{tag}_coa_map"]
    np.testing.assert_allclose(O.np_gaufilt3d(coa), G[f"{tag}_smoothed"], rtol=0, atol=1e-13)
    e, cov = O.np_covfit3d(coa, G[f"{tag}_spacing"])
    np.testing.assert_allclose(G[f"{tag}_ll"] + e, G[f"{tag}_cov_loc"], rtol=1e-12)
    np.testing.assert_allclose(np.diag(np.sqrt(np.abs(cov))), G[f"{tag}_cov_unc"], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("tag", TAGS)
def test_host_fits_match_reference(tag):
    """The host-side algebra of quakemigrate_b200.locate on the reference's own smoothed map."""
    from quakemigrate_b200 import locate
    coa, smoothed = G[f"{tag}_coa_map"], G[f"{tag}_smoothed"]
    lut = _LUT(coa.shape, G[f"{tag}_spacing"], G[f"{tag}_ll"])
    ijk = np.unravel_index(np.nanargmax(smoothed), smoothed.shape)
    idx, sig = locate.gaussian_fit(smoothed, ijk, np.nanmean(smoothed))
    np.testing.assert_allclose(lut.index2coord([list(idx)])[0], G[f"{tag}_gau_loc"], rtol=1e-9)
    np.testing.assert_allclose(sig * lut.node_spacing, G[f"{tag}_gau_unc"], rtol=1e-8)
    sp = locate.spline_location(coa, np.unravel_index(np.nanargmax(coa), coa.shape))
    np.testing.assert_allclose(lut.index2coord([list(sp)])[0], G[f"{tag}_spl_loc"], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_cuda_map_stage_matches_reference(tag):
    from quakemigrate_b200 import locate
    from quakemigrate_b200.core import Engine
    eng = Engine(0)
    coa = G[f"{tag}_coa_map"]
    lut = _LUT(coa.shape, G[f"{tag}_spacing"], G[f"{tag}_ll"])
    sm = eng.gaufilt3d(coa)
    np.testing.assert_allclose(sm, G[f"{tag}_smoothed"], rtol=0, atol=1e-12)
    ijk, mx, mean = eng.map_peak(coa)
    assert ijk == tuple(int(i) for i in np.unravel_index(np.nanargmax(coa), coa.shape))
    assert mx == coa.max() and mean == pytest.approx(coa.mean(), rel=1e-13)
    e, cov = eng.covfit3d(coa, lut.node_spacing)
    np.testing.assert_allclose(lut.ll_corner + e, G[f"{tag}_cov_loc"], rtol=1e-12)
    np.testing.assert_allclose(np.diag(np.sqrt(np.abs(cov))), G[f"{tag}_cov_unc"], rtol=1e-9, atol=1e-13)
    # the whole _calculate_location chain from an un-normalised marginal map
    res = locate.calculate_location(coa * 37.5, lut, eng, return_maps=True)
    np.testing.assert_allclose(res["coa_map"], coa, rtol=1e-15)
    np.testing.assert_allclose(res["spline"], G[f"{tag}_spl_loc"], rtol=1e-12)
    np.testing.assert_allclose(res["gaussian"][0], G[f"{tag}_gau_loc"], rtol=1e-8)
    np.testing.assert_allclose(res["gaussian"][1], G[f"{tag}_gau_unc"], rtol=1e-7)
    np.testing.assert_allclose(res["covariance"][0], G[f"{tag}_cov_loc"], rtol=1e-12)
    np.testing.assert_allclose(res["covariance"][1], G[f"{tag}_cov_unc"], rtol=1e-9, atol=1e-13)
    # windowed covariance (win option of scan.py:976-977)
    e_w, cov_w = eng.covfit3d(coa, lut.node_spacing, thresh=0.5, win=7)
    lo = [max(0, i - 3) for i in ijk]
    hi = [min(n, i + 4) for i, n in zip(ijk, coa.shape)]
    masked = np.zeros_like(coa)
    masked[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = coa[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
    e_r, cov_r = O.np_covfit3d(masked, lut.node_spacing, thresh=0.5)
    np.testing.assert_allclose(e_w, e_r, rtol=1e-12)
    np.testing.assert_allclose(cov_w, cov_r, rtol=1e-9, atol=1e-14)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_cuda_rbf_cubic_peak_matches_scipy(seed):
    """
    The dense evaluation of _splineloc's interpolant (scan.py:789-809) on the device against scipy's
    Rbf called exactly as the reference calls it: same fine-grid arg-max, peak value to rounding --
    on smooth peaks off the node centres (the locate case) and on rough random windows.
    """
    from scipy.interpolate import Rbf
    from quakemigrate_b200.core import Engine
    eng = Engine(0)
    rng = np.random.default_rng(seed)
    n, up = 5, 10
    g = np.arange(n) - 2.0
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    c = rng.uniform(-0.45, 0.45, 3)
    windows = [np.exp(-((X - c[0]) ** 2 / 2.5 + (Y - c[1]) ** 2 / 3.5 + (Z - c[2]) ** 2 / 1.7)),
               rng.uniform(0.0, 1.0, (n, n, n))]
    for trim in windows:
        axes = [np.linspace(0, n - 1, n)] * 3
        xo, yo, zo = np.meshgrid(*axes)
        rbf = Rbf(xo.flatten(), yo.flatten(), zo.flatten(), trim.flatten(), function="cubic")
        fine = [np.linspace(0, n - 1, (n - 1) * up + 1)] * 3
        xx, yy, zz = np.meshgrid(*fine)
        dense = rbf(xx.flatten(), yy.flatten(), zz.flatten()).reshape(xx.shape)
        want = tuple(int(i) for i in np.unravel_index(np.nanargmax(dense), dense.shape))
        got, val = eng.rbf_cubic_peak(rbf.nodes, n, up)
        assert val == pytest.approx(dense.max(), rel=1e-11)
        # a different fine node only if it ties with scipy's to rounding
        assert got == want or dense[got] == pytest.approx(dense.max(), rel=1e-12)
    eng.close()


def test_spline_window_with_a_device_style_peak_evaluator():
    """
    locate.spline_location_window hands the dense evaluation to ``engine.rbf_cubic_peak`` when the
    engine has it.  A numpy restatement of the kernel's index arithmetic (csrc/qmb_locate.cu,
    k_rbf_cubic: output (a, b, c) and centre (a', b', c') at the coordinates (b, a, c), linspace
    steps, first arg-max in C order) stands in for the device: both routes give the same location
    on smooth off-centre peaks, so the coordinate pairing of the reference's meshgrid calls
    (scan.py:777-804) is what the kernel reproduces.
    """
    from quakemigrate_b200 import locate

    class _Eng:
        calls = 0

        def rbf_cubic_peak(self, weights, n, upscale=10):
            type(self).calls += 1
            m = (n - 1) * upscale + 1
            step = (n - 1) / (m - 1)
            fine = np.arange(m) * step
            fine[-1] = n - 1
            a, b, c = np.meshgrid(fine, fine, fine, indexing="ij")      # output index (a, b, c)
            ca, cb, cc = [v.reshape(-1) for v in np.meshgrid(*[np.arange(n, dtype=float)] * 3, indexing="ij")]
            w = np.asarray(weights).reshape(-1)
            out = np.zeros(a.shape)
            for j in range(n ** 3):
                dx, dy, dz = b - cb[j], a - ca[j], c - cc[j]
                r = np.sqrt(dx * dx + dy * dy + dz * dz)
                out += r * r * r * w[j]
            flat = int(np.argmax(out))
            return tuple(int(i) for i in np.unravel_index(flat, out.shape)), float(out.reshape(-1)[flat])

    rng = np.random.default_rng(5)
    g = np.arange(5) - 2.0
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    for _ in range(4):
        c = rng.uniform(-0.45, 0.45, 3)
        trim = np.exp(-((X - c[0]) ** 2 / 2.5 + (Y - c[1]) ** 2 / 3.5 + (Z - c[2]) ** 2 / 1.7))
        lo, ijk = [10, 20, 30], [12, 22, 32]
        host = locate.spline_location_window(trim, lo, ijk)
        dev = locate.spline_location_window(trim, lo, ijk, engine=_Eng())
        np.testing.assert_array_equal(host, dev)
        np.testing.assert_allclose(host - np.array(ijk), c, atol=0.11)     # the fine grid is 0.1 node wide
    assert _Eng.calls == 4
