# -*- coding: utf-8 -*-
"""
Waveform pre-processing, host side: the Butterworth section design the product
carries (numpy only) against scipy's -- which is what obspy calls for the reference
(obspy/signal/filter.py: iirfilter(..., output="zpk") + zpk2sos) -- the oracle against the
committed golden vectors, and the small host pieces (NyquistException, upsample).
"""
import pathlib

import numpy as np
import pytest

from oracle import preprocess_oracle as PO
from quakemigrate_b200 import preprocess as pp

GOLD = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "golden_preprocess.npz")

CASES = [("bandpass", (2.0, 16.0), 50.0, 2), ("bandpass", (10.0, 124.0), 250.0, 4),
         ("bandpass", (0.1, 0.5), 100.0, 3), ("bandpass", (1.0, 20.0), 100.0, 1),
         ("lowpass", 50.0 / 2.000001, 100.0, 2), ("lowpass", 12.0, 100.0, 3), ("lowpass", 3.0, 100.0, 5),
         ("highpass", 2.0, 100.0, 2), ("highpass", 2.0, 100.0, 3), ("highpass", 20.0, 100.0, 1)]


@pytest.mark.parametrize("kind,freqs,df,corners", CASES)
def test_butter_sos_equals_scipy(kind, freqs, df, corners):
    """Same poles, same pairing, same section order, gain on the first section."""
    mine, ref = pp.butter_sos(kind, freqs, df, corners), PO.butter_sos(kind, freqs, df, corners)
    assert mine.shape == ref.shape
    np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=1e-15)


def test_butter_sos_matches_golden_sections():
    np.testing.assert_allclose(pp.butter_sos("bandpass", (2.0, 16.0), 50.0, 2), GOLD["sos_vt"], rtol=1e-12)
    np.testing.assert_allclose(pp.butter_sos("bandpass", (10.0, 124.0), 250.0, 4), GOLD["sos_ice"], rtol=1e-12)
    np.testing.assert_allclose(pp.butter_sos("lowpass", 50.0 / 2.000001, 100.0, 2), GOLD["sos_decimate"],
                               rtol=1e-12)


def test_butter_sos_rejects_bad_corners():
    with pytest.raises(ValueError):
        pp.butter_sos("bandpass", (2.0, 60.0), 100.0, 2)  # above Nyquist
    with pytest.raises(ValueError):
        pp.butter_sos("notch", 2.0, 100.0, 2)


def test_oracle_reproduces_golden():
    raw = GOLD["raw"]
    for name in ("vt", "ice", "low"):
        df, filt = float(GOLD[f"df_{name}"]), GOLD[f"filt_{name}"]
        got = np.stack([PO.pre_process(x, df, [filt[0], filt[1], int(filt[2])]) for x in raw])
        np.testing.assert_allclose(got, GOLD[f"pre_{name}"], rtol=0, atol=1e-9 * np.abs(GOLD[f"pre_{name}"]).max())
    got = np.stack([PO.decimate(x, 100.0, 50.0) for x in raw])
    assert got.shape == (3, 2000)
    np.testing.assert_allclose(got, GOLD["decimated"], rtol=0, atol=1e-9 * np.abs(GOLD["decimated"]).max())


def test_taper_window_shape():
    """Trace.taper(type="cosine", max_percentage=0.05): int(0.05 * npts) samples at either end."""
    for npts in (10, 25, 40, 41, 1000, 4001):
        w = PO.cosine_taper_window(npts)
        wlen = min(int(0.05 * npts), npts // 2)
        assert w.shape == (npts,) and np.allclose(w, w[::-1], atol=1e-15)
        assert (w[wlen:npts - wlen] == 1.0).all()
        if wlen:
            assert w[0] == 0.0 and w[-1] == 0.0 and (np.diff(w[:wlen]) > 0).all()
        if wlen > 1:
            assert abs(w[wlen - 1] - 1.0) < 1e-15


def test_pre_process_raises_nyquist_without_touching_the_gpu():
    with pytest.raises(pp.NyquistException):
        pp.pre_process(np.zeros((1, 100)), 50.0, [2.0, 25.0, 2])
    from quakemigrate_b200.onsets import STALTAOnset
    with pytest.raises(pp.NyquistException):
        STALTAOnset(sampling_rate=50, bandpass_filters={"P": [2, 16, 2], "S": [2, 25, 2]})


def test_upsample_is_the_reference_interpolation():
    """util.upsample (util.py:546-553): 40 Hz x 5 -> 200 Hz fence-post interpolation."""
    rng = np.random.default_rng(3)
    data = rng.integers(-1000, 1000, 17).astype(float)
    up = pp.upsample(data, 5)
    assert len(up) == (len(data) - 1) * 5 + 1
    np.testing.assert_array_equal(up[::5], data)
    np.testing.assert_allclose(up, np.interp(np.arange(len(up)) / 5.0, np.arange(len(data)), data), atol=1e-12)
