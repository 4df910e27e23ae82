# -*- coding: utf-8 -*-
"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's waveform pre-processing
(SURVEY 8f rank 4).  Imported by tests/ (and tests/golden/make_golden_preprocess.py); never by the
product.

What the reference does (``quakemigrate/signal/onsets/stalta.py:137-211`` ``pre_process``;
``quakemigrate/util.py:479-520`` ``decimate``) is a chain of obspy calls:

    detrend("linear") -> detrend("constant"/"demean") -> taper(type="cosine", max_percentage=0.05)
    -> filter("bandpass", freqmin, freqmax, corners, zerophase=True)          (pre_process)
    -> filter("lowpass", freq=sr/2.000001, corners=2, zerophase=True) -> data[::factor]   (decimate)

obspy is a third-party dependency that is NOT present in /root/reference nor installable here
(requirements: ``obspy>=1.3``, reference ``pyproject.toml``); its routines are thin wrappers over
scipy.signal, which IS installed, so the arithmetic below is scipy's own:

* ``Trace.detrend("linear")`` / ``("constant")``  = ``scipy.signal.detrend(data, type=...)``
  (obspy/core/trace.py ``detrend``);
* ``Trace.filter("bandpass", zerophase=True)`` = ``iirfilter(corners, [f1/fe, f2/fe], btype="band",
  ftype="butter", output="zpk")`` -> ``zpk2sos`` -> ``sosfilt`` forward, then ``sosfilt`` on the
  reversed result, reversed again (obspy/signal/filter.py ``bandpass`` / ``lowpass``);
* ``Trace.taper(type="cosine", max_percentage=p)``: ``wlen = min(int(p * npts), int(npts / 2))``
  samples at each end are multiplied by the two halves of
  ``obspy.signal.invsim.cosine_taper(2 * wlen + 1, p=1.0)``, i.e. ``0.5 * (1 - cos(pi * k / (wlen - 1)))``
  for ``k = 0 .. wlen-1`` on the left and its mirror image on the right (restated from the
  published obspy source; PARITY OF THE TAPER IS UNPINNED here -- no obspy to run).

Everything else is pinned against scipy run in this container; golden vectors in
``tests/golden/golden_preprocess.npz`` (made by ``tests/golden/make_golden_preprocess.py``).
"""

from __future__ import annotations

import numpy as np
from scipy.signal import detrend as _sp_detrend
from scipy.signal import hilbert as _sp_hilbert
from scipy.signal import iirfilter, sosfilt, zpk2sos


def cosine_taper_window(npts: int, max_percentage: float = 0.05) -> np.ndarray:
    """Full-length taper ``Trace.taper(type="cosine", max_percentage=...)`` multiplies by."""
    wlen = min(int(max_percentage * npts), int(npts / 2))
    m = 2 * wlen if 2 * wlen == npts else 2 * wlen + 1
    # obspy.signal.invsim.cosine_taper(m, p=1.0)
    frac = int(m * 1.0 / 2.0)
    idx1, idx2, idx3, idx4 = 0, frac - 1, m - frac, m - 1
    if idx1 == idx2:
        idx2 += 1
    if idx3 == idx4:
        idx3 -= 1
    sides = np.zeros(m)
    if m > 0:
        sides[idx1:idx2 + 1] = 0.5 * (1.0 - np.cos(np.pi * (np.arange(idx1, idx2 + 1) - float(idx1))
                                                   / (idx2 - idx1)))
        sides[idx2 + 1:idx3] = 1.0
        sides[idx3:idx4 + 1] = 0.5 * (1.0 + np.cos(np.pi * (float(idx3) - np.arange(idx3, idx4 + 1))
                                                   / (idx4 - idx3)))
    return np.hstack((sides[:wlen], np.ones(npts - 2 * wlen), sides[len(sides) - wlen:]))


def butter_sos(kind: str, freqs, df: float, corners: int) -> np.ndarray:
    """obspy.signal.filter.bandpass / lowpass / highpass coefficient design (scipy)."""
    fe = 0.5 * df
    if kind == "bandpass":
        wn = [freqs[0] / fe, freqs[1] / fe]
        btype = "band"
    else:
        wn = freqs / fe
        btype = kind
    z, p, k = iirfilter(corners, wn, btype=btype, ftype="butter", output="zpk")
    return zpk2sos(z, p, k)


def sosfilt_zerophase(sos: np.ndarray, data: np.ndarray) -> np.ndarray:
    firstpass = sosfilt(sos, data)
    return sosfilt(sos, firstpass[::-1])[::-1]


def detrend_taper(x: np.ndarray, max_percentage: float = 0.05) -> np.ndarray:
    x = np.require(x, dtype=np.float64)
    x = _sp_detrend(x, type="linear")
    x = _sp_detrend(x, type="constant")
    return x * cosine_taper_window(len(x), max_percentage)


def pre_process(x: np.ndarray, df: float, filter_, sos: np.ndarray | None = None) -> np.ndarray:
    """stalta.py:203-209 for one trace already at the target sampling rate."""
    lowcut, highcut, order = filter_
    if sos is None:
        sos = butter_sos("bandpass", (lowcut, highcut), df, order)
    return sosfilt_zerophase(sos, detrend_taper(x))


def decimate(x: np.ndarray, df: float, sampling_rate: float, sos: np.ndarray | None = None) -> np.ndarray:
    """util.py:479-520 (``Trace.decimate(factor, no_filter=True)`` is ``data[::factor]``)."""
    if sos is None:
        sos = butter_sos("lowpass", float(sampling_rate) / 2.000001, df, 2)
    y = sosfilt_zerophase(sos, detrend_taper(x))
    return y[:: int(df / sampling_rate)]


def envelope(x: np.ndarray) -> np.ndarray:
    """stalta.py:520-523: ``np.abs(hilbert(tr.data))`` (scipy.signal.hilbert, FFT length = len(x))."""
    return np.abs(_sp_hilbert(np.asarray(x, dtype=np.float64), axis=-1))
