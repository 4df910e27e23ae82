# -*- coding: utf-8 -*-
"""
Waveform pre-processing on the GPU -- what the reference does with obspy before
the onset functions:

* :func:`pre_process` -- ``quakemigrate/signal/onsets/stalta.py:137-211``: detrend (linear, then
  constant), 5 % cosine taper, zero-phase Butterworth band-pass ``[lowcut, highcut, corners]``;
  raises :class:`NyquistException` like stalta.py:198-199.
* :func:`decimate` -- ``quakemigrate/util.py:479-520``: the same with a 2-corner low-pass at
  ``sampling_rate / 2.000001``, then every ``factor``-th sample.
* :func:`upsample` -- ``quakemigrate/util.py:523-604`` (linear interpolation; a host gather, kept
  in numpy: it is index arithmetic, not a filter).
* :func:`butter_sos` -- the coefficient design obspy delegates to scipy
  (``iirfilter(..., ftype="butter", output="zpk")`` + ``zpk2sos``; obspy/signal/filter.py), restated
  in numpy so that the product does not import scipy: analogue Butterworth prototype, band / low /
  high-pass transform, bilinear transform, pole/zero pairing "nearest", worst pole last, gain on
  the first section.  tests/test_preprocess.py checks the sections against scipy's.

The traces are filtered by ``libqmb200.so`` (``qmb_waveforms_filter``: chunk-parallel IIR, see
``csrc/qmb_preprocess.cu``) and can stay on the device for the STA/LTA stage
(:meth:`quakemigrate_b200.core.Engine.onsets_from_device_waveforms`).  No CPU fallback.
"""

from __future__ import annotations

import numpy as np

__all__ = ["butter_sos", "pre_process", "decimate", "upsample", "cosine_taper_window", "NyquistException"]


class NyquistException(Exception):
    """``quakemigrate.util.NyquistException``: filter corner at or above the Nyquist frequency."""

    def __init__(self, freqmax, f_nyquist, tr_id):
        super().__init__(
            f"    NyquistException: Selected bandpass_highcut {freqmax} Hz is at or above the "
            f"Nyquist frequency ({f_nyquist} Hz) for trace {tr_id}."
        )


# ------------------------------------------------------------------------------------------------
# Butterworth second-order sections
# ------------------------------------------------------------------------------------------------
def _section(zeros, poles, gain):
    return np.concatenate([gain * np.real(np.poly(zeros)), np.real(np.poly(poles))])


def butter_sos(kind: str, freqs, df: float, corners: int) -> np.ndarray:
    """
    Second-order sections ``(n_sections, 6)`` of the digital Butterworth filter obspy builds for
    ``Trace.filter(kind, ..., corners=corners)`` at sampling rate ``df``.

    kind: ``"bandpass"`` (``freqs = (freqmin, freqmax)``), ``"lowpass"`` or ``"highpass"``
    (``freqs`` = corner in Hz).
    """
    n = int(corners)
    if n < 1:
        raise ValueError("corners must be >= 1")
    fe = 0.5 * df
    # analogue prototype (scipy.signal.buttap)
    p = -np.exp(1j * np.pi * np.arange(-n + 1, n, 2) / (2 * n))
    z = np.array([], dtype=complex)
    k = 1.0
    fs = 2.0
    if kind == "bandpass":
        wn = np.array([freqs[0] / fe, freqs[1] / fe], dtype=float)
    elif kind in ("lowpass", "highpass"):
        wn = np.array([float(freqs) / fe])
    else:
        raise ValueError(f"unknown filter kind {kind!r}")
    if np.any(wn <= 0) or np.any(wn >= 1):
        raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
    warped = 2 * fs * np.tan(np.pi * wn / fs)
    degree = len(p) - len(z)
    if kind == "lowpass":  # lp2lp_zpk
        z, p, k = z * warped[0], p * warped[0], k * warped[0] ** degree
    elif kind == "highpass":  # lp2hp_zpk
        k = k * np.real(np.prod(-z) / np.prod(-p))
        z, p = warped[0] / z, warped[0] / p
        z = np.append(z, np.zeros(degree))
    else:  # lp2bp_zpk
        bw, wo = warped[1] - warped[0], np.sqrt(warped[0] * warped[1])
        z_lp, p_lp = z * bw / 2, p * bw / 2
        z_lp, p_lp = z_lp.astype(complex), p_lp.astype(complex)
        z = np.concatenate((z_lp + np.sqrt(z_lp ** 2 - wo ** 2), z_lp - np.sqrt(z_lp ** 2 - wo ** 2)))
        p = np.concatenate((p_lp + np.sqrt(p_lp ** 2 - wo ** 2), p_lp - np.sqrt(p_lp ** 2 - wo ** 2)))
        z = np.append(z, np.zeros(degree))
        k = k * bw ** degree
    # bilinear_zpk
    degree = len(p) - len(z)
    fs2 = 2.0 * fs
    z_z, p_z = (fs2 + z) / (fs2 - z), (fs2 + p) / (fs2 - p)
    z_z = np.append(z_z, -np.ones(degree))
    k_z = k * np.real(np.prod(fs2 - z) / np.prod(fs2 - p))

    # zpk2sos, pairing "nearest": poles and zeros padded with 0 to an even count, one
    # representative per conjugate pair, worst pole (closest to the unit circle) served first with
    # the two nearest remaining zeros (every Butterworth zero is real: +1 / -1, or the padding 0)
    cplx = p_z[np.imag(p_z) > 1e-12 * np.abs(p_z)]
    real = list(np.real(p_z[np.abs(np.imag(p_z)) <= 1e-12 * np.abs(p_z)]))
    zeros = list(np.real(z_z))
    if len(p_z) % 2:
        real.append(0.0)
        zeros.append(0.0)
    real.sort(key=lambda r: abs(1 - abs(r)))
    pole_groups = [(pc, np.conj(pc)) for pc in cplx]
    while real:
        pole_groups.append((real.pop(0), real.pop(0)))
    pole_groups.sort(key=lambda g: abs(1 - abs(g[0])))
    sections = []
    for group in pole_groups:
        take = []
        for _ in range(2):
            j = int(np.argmin([abs(group[0] - zz) for zz in zeros]))
            take.append(zeros.pop(j))
        sections.append((take, list(group)))
    sections = sections[::-1]  # worst last
    sos = np.zeros((len(sections), 6))
    for i, (zz, pp) in enumerate(sections):
        sos[i] = _section(zz, pp, k_z if i == 0 else 1.0)
    return sos


def cosine_taper_window(npts: int, max_percentage: float = 0.05) -> np.ndarray:
    """
    The window ``Trace.taper(type="cosine", max_percentage=p)`` of obspy multiplies a trace by
    (what stalta.py:207 and :448 apply).  obspy takes ``w = min(int(p * npts), npts // 2)`` samples
    per side from ``cosine_taper(2 w + 1, p=1.0)`` (``2 w`` points when they span the whole trace):
    a half cosine that is 0 on the first sample and reaches 1 on sample ``w - 1``; everything
    between the two sides is 1.  Host form of the device taper in ``csrc/qmb_preprocess.cu`` (used
    for the segments of gappy channels); the device taper is pinned end to end against obspy's own
    output by tests/test_iceland_locate.py, this one against the device and the oracle.
    """
    npts = int(npts)
    w = min(int(max_percentage * npts), npts // 2)
    win = np.ones(npts)
    if w > 0:
        m = 2 * w if 2 * w == npts else 2 * w + 1      # length of the generating window
        rise_end = max(m // 2 - 1, 1)                   # index where the rising half reaches 1 (>= 1)
        fall_start = m - m // 2
        if fall_start == m - 1:                         # degenerate one-interval falling half
            fall_start -= 1
        k = np.arange(w)
        win[:w] = np.where(k <= rise_end, 0.5 * (1.0 - np.cos(np.pi * k / rise_end)), 1.0)
        j = np.arange(m - w, m)                         # last w samples of the generating window
        fall = 0.5 * (1.0 + np.cos(np.pi * (fall_start - j) / (m - 1 - fall_start)))
        win[npts - w:] = np.where(j >= fall_start, fall, 1.0)
    return win


# ------------------------------------------------------------------------------------------------
# the obspy chains
# ------------------------------------------------------------------------------------------------
def _engine(engine):
    if engine is None:
        from .core import Engine
        engine = Engine(0)
    return engine


def pre_process(waveforms, sampling_rate, filter_, engine=None, return_device=False):
    """
    ``pre_process`` (stalta.py:137-211) for an ``(n_channels, n)`` block of traces that are already
    at ``sampling_rate`` (the resampling step is :func:`decimate` / :func:`upsample`).
    ``filter_ = [lowcut, highcut, corners]``.  Returns the filtered block (float64), or the
    device-resident :class:`~quakemigrate_b200.core.DeviceWaveforms` when ``return_device``.
    """
    lowcut, highcut, order = filter_
    if highcut >= 0.5 * sampling_rate:
        raise NyquistException(highcut, 0.5 * sampling_rate, "")
    engine = _engine(engine)
    sos = butter_sos("bandpass", (lowcut, highcut), sampling_rate, order)
    raw = engine.upload_waveforms(waveforms)
    try:
        out = engine.filter_waveforms(raw, sos, detrend=True, taper_fraction=0.05, zerophase=True)
    finally:
        raw.free()
    if return_device:
        return out
    try:
        return out.fetch()
    finally:
        out.free()


def decimate(waveforms, trace_sampling_rate, sampling_rate, engine=None):
    """``util.decimate`` (util.py:479-520) for an ``(n_channels, n)`` block: returns ``(C, ceil(n/f))``."""
    factor = int(trace_sampling_rate / sampling_rate)
    engine = _engine(engine)
    sos = butter_sos("lowpass", float(sampling_rate) / 2.000001, trace_sampling_rate, 2)
    raw = engine.upload_waveforms(waveforms)
    try:
        out = engine.filter_waveforms(raw, sos, detrend=True, taper_fraction=0.05, zerophase=True,
                                      decimate=factor)
    finally:
        raw.free()
    try:
        return out.fetch()
    finally:
        out.free()


def upsample(data, upfactor):
    """The interpolation of ``util.upsample`` (util.py:546-553): fence-post linear interpolation."""
    data = np.asarray(data)
    dnew = np.zeros((len(data) - 1) * upfactor + 1)
    dnew[::upfactor] = data
    for i in range(1, upfactor):
        dnew[i::upfactor] = float(i) / upfactor * data[1:] + float(upfactor - i) / upfactor * data[:-1]
    return dnew
