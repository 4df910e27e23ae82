# -*- coding: utf-8 -*-
"""
``QuakeScan._calculate_location`` (``quakemigrate/signal/scan.py:696-733``) with the grid-sized
work on the GPU: given the marginal-window sum of the coalescence volume
(``coa_map = sum_t map4d[..., i0:i1]``, produced on the device by ``Engine.locate_window``) it
returns the three location estimates of the reference:

* spline   -- ``_splineloc`` (scan.py:736-841): cubic RBF on the 5^3 window around the peak,
              upsampled x10 (small; host, scipy -- the same library call the reference makes);
* gaussian -- ``_gaufilt3d`` (scan.py:1008-1043) on the device, then ``_gaufit3d``
              (scan.py:844-936): log-quadratic least squares on the 7^3 window (host, 343x10);
* covariance -- ``_covfit3d`` (scan.py:939-1005): weighted moments of the nodes above the 0.90
              level, reduced on the device.

Coordinates: ``lut.index2coord`` maps fractional grid indices to the caller's coordinates exactly
like the reference (``ll_corner + idx * node_spacing`` followed by the optional projection).
"""

from __future__ import annotations

import logging

import numpy as np


def _window(shape, ijk, win):
    """Slices of the +-(win-1)//2 window around ``ijk`` clipped to the grid (scan.py:1045-1070)."""
    w2 = (win - 1) // 2
    lo = [int(np.clip(i - w2, 0, n)) for i, n in zip(ijk, shape)]
    hi = [int(np.clip(i + w2 + 1, 0, n)) for i, n in zip(ijk, shape)]
    return lo, hi


def spline_location_window(trim: np.ndarray, lo, ijk, win: int = 5, upscale: int = 10, engine=None):
    """
    ``_splineloc`` (scan.py:736-841) on the window ``trim`` = ``coa_map[lo : lo + trim.shape]``:
    returns the fractional grid index of the interpolated peak, or the gridded peak when the
    window touches the grid edge or the result leaves the window.  With an ``engine`` that has
    ``rbf_cubic_peak`` the interpolation weights come from scipy's ``Rbf`` (a 125 x 125 solve) and
    the dense evaluation on the 41^3 fine grid plus its arg-max run on the device.
    """
    from scipy.interpolate import Rbf

    ijk = np.asarray(ijk)
    ext = trim.shape
    if not (ext[0] == ext[1] == ext[2]):
        logging.info("\t !!!! Spline error: interpolation window crosses edge of grid !!!!")
        return ijk.astype(float)
    axes = [np.linspace(0, n - 1, n) for n in trim.shape]
    # NB: the reference builds its sample grid with numpy.meshgrid's default 'xy' indexing
    # (scan.py:777, :797) and flattens the data in C order; reproduce exactly that pairing.
    xog, yog, zog = np.meshgrid(*axes)
    rbf = Rbf(xog.flatten(), yog.flatten(), zog.flatten(), trim.flatten(), function="cubic")
    if engine is not None and hasattr(engine, "rbf_cubic_peak"):
        fine_idx, _ = engine.rbf_cubic_peak(rbf.nodes, ext[0], upscale)
    else:
        fine = [np.linspace(0, n - 1, (n - 1) * upscale + 1) for n in trim.shape]
        xxg, yyg, zzg = np.meshgrid(*fine)
        interp = rbf(xxg.flatten(), yyg.flatten(), zzg.flatten()).reshape(xxg.shape)
        fine_idx = np.unravel_index(np.nanargmax(interp), interp.shape)
    peak = np.array(fine_idx) / upscale + np.array(lo)
    w2 = (win - 1) // 2
    if np.any(np.abs(ijk - peak) > w2):
        logging.info("\t !!!! Spline error: location outside interpolation window !!!!")
        return ijk.astype(float)
    return peak


def spline_location(coa_map: np.ndarray, ijk, win: int = 5, upscale: int = 10):
    """``_splineloc`` on a full host map: cut the window, then :func:`spline_location_window`."""
    lo, hi = _window(coa_map.shape, ijk, win)
    return spline_location_window(coa_map[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]], lo, ijk, win, upscale)


def gaussian_fit_window(sub: np.ndarray, lo, ijk, mean: float, thresh: float = 0.0):
    """
    ``_gaufit3d`` (scan.py:844-936) on the window ``sub`` = ``smoothed[lo : lo + sub.shape]`` around
    the peak ``ijk``: fit ``-log(map - mean) = quadratic(x, y, z)`` over the window nodes above
    ``thresh``; returns (fractional grid index of the expectation, 1-sigma in nodes).
    """
    hi = [l + n for l, n in zip(lo, sub.shape)]
    keep = sub > thresh                                             # scan.py:873 (on the raw map)
    gx, gy, gz = np.meshgrid(*[np.arange(l, h) - c for l, h, c in zip(lo, hi, ijk)], indexing="ij")
    x, y, z = gx[keep].astype(float), gy[keep].astype(float), gz[keep].astype(float)
    vals = (sub - mean)[keep]                                       # scan.py:878
    design = np.c_[x * x, y * y, z * z, x * y, x * z, y * z, x, y, z, np.ones(len(x))].T
    target = -np.log(np.clip(vals.astype(np.float64), 1e-300, np.inf))
    coef = np.matmul(target, np.linalg.pinv(design))
    hess = -np.array([[2 * coef[0], coef[3], coef[4]],
                      [coef[3], 2 * coef[1], coef[5]],
                      [coef[4], coef[5], 2 * coef[2]]])
    loc = np.matmul(np.linalg.inv(hess), coef[6:9])
    quad = np.array([[coef[0], coef[3] / 2, coef[4] / 2],
                     [coef[3] / 2, coef[1], coef[5] / 2],
                     [coef[4] / 2, coef[5] / 2, coef[2]]])
    egv, _ = np.linalg.eig(quad)
    sigma_nodes = np.sqrt(0.5 / np.clip(np.abs(egv), 1e-10, np.inf)) / 2
    return loc + np.asarray(ijk), sigma_nodes


def gaussian_fit(smoothed: np.ndarray, ijk, mean: float, thresh: float = 0.0, win: int = 7):
    """``_gaufit3d`` on a full host map: cut the window, then :func:`gaussian_fit_window`."""
    lo, hi = _window(smoothed.shape, ijk, win)
    return gaussian_fit_window(smoothed[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]], lo, ijk, mean, thresh)


def _finish(lut, spline_idx, gau_idx, gau_sigma, expect, cov, ijk):
    spline_loc = lut.index2coord([list(spline_idx)])[0]
    gau_loc = lut.index2coord([list(gau_idx)])[0]
    gau_unc = gau_sigma * np.asarray(lut.node_spacing)
    cov_xyz = np.asarray(lut.ll_corner) + expect                    # scan.py:1001
    cov_loc = lut.grid2coord(np.atleast_2d(cov_xyz))[0] if getattr(lut, "grid2coord", None) else cov_xyz
    cov_unc = np.diag(np.sqrt(np.abs(cov)))                         # scan.py:1003
    return {"spline": spline_loc, "gaussian": (gau_loc, gau_unc), "covariance": (cov_loc, cov_unc),
            "peak_index": ijk}


def calculate_location_device(coa_map_sum, lut, return_maps: bool = False, engine_timer=None) -> dict:
    """
    ``_calculate_location`` (scan.py:696-733) on a :class:`~quakemigrate_b200.core.DeviceMap`
    holding the un-normalised marginal map: the map is normalised, smoothed and reduced on the
    device; only the peak (32 B), the 5^3 spline window, the 7^3 Gaussian-fit window and the 12
    covariance moments cross PCIe (< 5 KB per event; ``pcie_bytes`` in the result).
    ``engine_timer`` (benchmarks): called once the last device operation of the event has
    finished, before the host-side fits.
    """
    # device work first (one stream, a few small read-backs), host algebra afterwards
    coa = coa_map_sum.normalise()                                   # scan.py:720-721
    ijk, _, _ = coa.peak()
    lo, hi = _window(coa.shape, ijk, 5)
    trim = coa.window(lo, hi)
    smoothed = coa.gaufilt()
    s_ijk, _, s_mean = smoothed.peak()
    lo7, hi7 = _window(coa.shape, s_ijk, 7)
    sub = smoothed.window(lo7, hi7)
    expect, cov = coa.covfit(lut.node_spacing, thresh=0.90)
    if engine_timer is not None:
        engine_timer()
    spline_idx = spline_location_window(trim, lo, ijk, engine=getattr(coa_map_sum, "engine", None))
    gau_idx, gau_sigma = gaussian_fit_window(sub, lo7, s_ijk, s_mean)
    out = _finish(lut, spline_idx, gau_idx, gau_sigma, expect, cov, ijk)
    out["pcie_bytes"] = int(2 * 32 + trim.nbytes + sub.nbytes + 12 * 8 + 2 * 7 * 8 + trim.nbytes + 16)
    if return_maps:
        out["coa_map"], out["smoothed"] = coa.fetch(), smoothed.fetch()
    smoothed.free()
    return out


def calculate_location(coa_map_sum, lut, engine, return_maps: bool = False) -> dict:
    """
    ``_calculate_location`` (scan.py:696-733).  ``coa_map_sum`` is the un-normalised marginal map:
    a device map (the product path, see :func:`calculate_location_device`) or a host array, which
    is uploaded once when the engine can keep maps on the device.
    Returns a dict with ``spline`` location, ``gaussian`` (location, uncertainty), ``covariance``
    (location, uncertainty) and, with ``return_maps``, ``coa_map`` (normalised) and ``smoothed``.
    """
    if hasattr(coa_map_sum, "normalise"):
        return calculate_location_device(coa_map_sum, lut, return_maps)
    if hasattr(engine, "upload_map"):
        dev = engine.upload_map(coa_map_sum)
        try:
            return calculate_location_device(dev, lut, return_maps)
        finally:
            dev.free()
    coa_map = coa_map_sum / np.nanmax(coa_map_sum)
    ijk, _, _ = engine.map_peak(coa_map)
    spline_idx = spline_location(coa_map, ijk)
    smoothed = engine.gaufilt3d(coa_map)
    s_ijk, _, s_mean = engine.map_peak(smoothed)
    gau_idx, gau_sigma = gaussian_fit(smoothed, s_ijk, s_mean)
    expect, cov = engine.covfit3d(coa_map, lut.node_spacing, thresh=0.90)
    out = _finish(lut, spline_idx, gau_idx, gau_sigma, expect, cov, ijk)
    if return_maps:
        out["coa_map"], out["smoothed"] = coa_map, smoothed
    return out
