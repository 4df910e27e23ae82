# -*- coding: utf-8 -*-
"""
Waveform pre-processing on the device: ``qmb_waveforms_filter`` /
``qmb_preprocess`` against the oracle (scipy.signal, which is what obspy runs for the reference:
``quakemigrate/signal/onsets/stalta.py:203-209``, ``quakemigrate/util.py:499-518``) and against the
committed golden vectors.  float64 throughout; tolerance 1e-9 of the trace's peak (the chunked
recurrence re-associates the state propagation: rounding-level differences, amplified by poles
close to the unit circle to ~1e-11 for narrow low-frequency bands).
"""
import pathlib

import numpy as np
import pytest

from oracle import preprocess_oracle as PO
from quakemigrate_b200 import preprocess as pp

pytestmark = pytest.mark.gpu
GOLD = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "golden_preprocess.npz")
RTOL = 1e-9


def _close(got, want, rtol=RTOL):
    assert got.shape == want.shape
    scale = np.abs(want).max(axis=-1, keepdims=True)
    scale[scale == 0] = 1.0
    err = np.abs(got - want) / scale
    assert err.max() < rtol, f"max relative error {err.max():.3e}"


def _traces(rng, c, n, trend=True):
    x = rng.standard_normal((c, n)) * rng.uniform(1.0, 1e4, (c, 1))
    if trend:
        x += rng.uniform(-1e5, 1e5, (c, 1)) + rng.uniform(-3.0, 3.0, (c, 1)) * np.arange(n)
    return x


@pytest.fixture(scope="module")
def eng():
    from quakemigrate_b200.core import Engine
    with Engine(0) as e:
        yield e


def test_golden_pre_process(eng):
    raw = GOLD["raw"]
    for name in ("vt", "ice", "low"):
        df, filt = float(GOLD[f"df_{name}"]), GOLD[f"filt_{name}"]
        got = pp.pre_process(raw, df, [filt[0], filt[1], int(filt[2])], engine=eng)
        _close(got, GOLD[f"pre_{name}"])
    _close(pp.decimate(raw, 100.0, 50.0, engine=eng), GOLD["decimated"])
    _close(eng.preprocess(raw, None), GOLD["detrend_taper"], rtol=1e-11)


@pytest.mark.parametrize("n", [3, 19, 20, 33, 40, 1000, 4095, 4096, 4097, 8191, 8192, 8193, 20000, 70001])
def test_lengths_around_chunk_and_tile_edges(eng, n):
    rng = np.random.default_rng(n)
    x = _traces(rng, 3, n)
    sos = PO.butter_sos("bandpass", (2.0, 16.0), 100.0, 2)
    want = np.stack([PO.sosfilt_zerophase(sos, PO.detrend_taper(t)) for t in x])
    _close(eng.preprocess(x, sos), want)


@pytest.mark.parametrize("corners,band", [(1, (1.0, 20.0)), (2, (2.0, 16.0)), (3, (0.5, 5.0)), (4, (10.0, 45.0))])
@pytest.mark.parametrize("zerophase", [True, False])
def test_section_counts_and_single_pass(eng, corners, band, zerophase):
    from scipy.signal import sosfilt
    rng = np.random.default_rng(corners)
    x = _traces(rng, 5, 30011)
    sos = PO.butter_sos("bandpass", band, 100.0, corners)
    assert len(sos) == corners
    pre = [PO.detrend_taper(t) for t in x]
    want = np.stack([PO.sosfilt_zerophase(sos, t) if zerophase else sosfilt(sos, t) for t in pre])
    _close(eng.preprocess(x, sos, zerophase=zerophase), want)


def test_too_many_sections_is_refused(eng):
    from quakemigrate_b200.core import QmbError
    sos = PO.butter_sos("bandpass", (2.0, 16.0), 100.0, 5)
    with pytest.raises(QmbError) as e:
        eng.preprocess(np.zeros((1, 100)), sos)
    assert e.value.status == 3


@pytest.mark.parametrize("factor", [2, 4, 5])
def test_decimate(eng, factor):
    """util.decimate (util.py:479-520): 100 Hz -> 100/factor Hz."""
    rng = np.random.default_rng(factor)
    x = _traces(rng, 4, 12345)
    want = np.stack([PO.decimate(t, 100.0, 100.0 / factor) for t in x])
    got = pp.decimate(x, 100.0, 100.0 / factor, engine=eng)
    assert got.shape == (4, -(-12345 // factor))
    _close(got, want)


def test_options_detrend_taper_filter_independently(eng):
    from scipy.signal import sosfilt
    rng = np.random.default_rng(5)
    x = _traces(rng, 2, 5000)
    sos = PO.butter_sos("lowpass", 10.0, 100.0, 2)
    _close(eng.preprocess(x, sos, detrend=False, taper_fraction=0.0, zerophase=False),
           np.stack([sosfilt(sos, t) for t in x]))
    _close(eng.preprocess(x, None, detrend=False, taper_fraction=0.05),
           np.stack([t * PO.cosine_taper_window(len(t)) for t in x]), rtol=1e-14)
    from scipy.signal import detrend
    _close(eng.preprocess(x, None, detrend=True, taper_fraction=0.0),
           np.stack([detrend(detrend(t, type="linear"), type="constant") for t in x]), rtol=1e-11)


def test_channel_selection_and_output_offset(eng):
    rng = np.random.default_rng(6)
    x = _traces(rng, 6, 9000)
    sos_a = PO.butter_sos("bandpass", (2.0, 16.0), 100.0, 2)
    sos_b = PO.butter_sos("bandpass", (1.0, 8.0), 100.0, 3)
    raw = eng.upload_waveforms(x)
    out = eng.alloc_waveforms(6, 9000)
    eng.filter_waveforms(raw, sos_a, channels=[0, 2, 3], out=out, out_chan0=0)
    eng.filter_waveforms(raw, sos_b, channels=[5, 1], out=out, out_chan0=3)
    got = out.fetch()
    want = [PO.sosfilt_zerophase(sos_a, PO.detrend_taper(x[c])) for c in (0, 2, 3)]
    want += [PO.sosfilt_zerophase(sos_b, PO.detrend_taper(x[c])) for c in (5, 1)]
    _close(got[:5], np.stack(want))
    assert (got[5] == 0).all()  # untouched channel of a fresh block
    _close(out.fetch(3, 2), np.stack(want[3:]))
    raw.free()
    out.free()


def test_full_size_block_linearity_and_spot_parity(eng):
    """cfg3-sized block (96 traces x 380 000 samples): linearity of the whole chain, and four
    traces against the oracle at full length."""
    rng = np.random.default_rng(7)
    c, n = 96, 380_000
    a = _traces(rng, c, n)
    b = _traces(rng, c, n)
    sos = pp.butter_sos("bandpass", (2.0, 16.0), 100.0, 2)
    fa, fb = eng.preprocess(a, sos), eng.preprocess(b, sos)
    fab = eng.preprocess(2.0 * a - 0.5 * b, sos)
    _close(fab, 2.0 * fa - 0.5 * fb, rtol=1e-9)
    for k in (0, 31, 64, 95):
        _close(fa[k:k + 1], PO.pre_process(a[k], 100.0, [2.0, 16.0, 2])[None, :])


def test_stalta_onset_with_bandpass_filters_matches_the_reference_chain(eng):
    """STALTAOnset(bandpass_filters=...).calculate_onsets == pre_process -> _onset of the reference
    (stalta.py:420-470, :516-548) built from the oracle's pieces."""
    from oracle import qm_oracle as O
    from quakemigrate_b200.onsets import STALTAOnset, WaveformData

    rng = np.random.default_rng(8)
    sr, n = 100, 6000
    stations = [f"ST{k}" for k in range(4)]
    channels = {st: {c: _traces(rng, 1, n)[0] for c in "ZNE"} for st in stations}
    del channels["ST1"]["E"]
    filters = {"P": [2.0, 16.0, 2], "S": [1.0, 10.0, 3]}
    onset = STALTAOnset(engine=eng, sampling_rate=sr, position="classic", bandpass_filters=filters)
    onsets, onset_data = onset.calculate_onsets(WaveformData(channels, 0.0, sr))
    nsta, nlta = 21, 101
    want = []
    for phase, comps in (("P", "Z"), ("S", "NE")):
        for st in stations:
            have = [c for c in comps if c in channels[st]]
            filt = [PO.pre_process(channels[st][c], sr, filters[phase]) for c in have]
            ons = np.array([O.overlapping_sta_lta(f ** 2, nsta, nlta) for f in filt])
            want.append(np.clip(np.sqrt(np.sum(ons ** 2, axis=0) / len(ons)), 0.4, np.inf))
            got_f = onset_data.filtered_waveforms.trace(st, have[0], phase)
            _close(got_f[None, :], filt[0][None, :])
    want = np.stack(want)
    assert onsets.shape == want.shape
    np.testing.assert_allclose(onsets, want, rtol=1e-6)
    assert onset_data.device_onsets is not None and sum(onset_data.availability.values()) == 8


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 17, 64, 1001, 4096, 4097, 30011, 65536, 137460, 409260])
def test_envelope_matches_scipy_hilbert(eng, n):
    """
    |analytic signal| for even / odd / prime / composite / power-of-two lengths (137460 and 409260 are
    cfg2's and cfg3's row lengths): the library's own Bluestein + Stockham transforms against scipy's
    exact-length FFT.
    """
    rng = np.random.default_rng(n)
    x = _traces(rng, 3, n, trend=False)
    _close(eng.envelope(x), PO.envelope(x), rtol=1e-11)


@pytest.mark.parametrize("transform", ["env", "env_squared"])
def test_onsets_with_envelope_transforms(eng, transform):
    """stalta.py:520-523 + :525-548 with the envelope transforms (dike_intrusion_detect.py:47)."""
    from oracle import qm_oracle as O
    rng = np.random.default_rng(11)
    n, nsta, nlta = 9000, 11, 51
    wf = _traces(rng, 5, n, trend=False)
    rows = [[0], [1, 2], [3, 4]]
    dev, got = eng.onsets_from_waveforms(wf, rows, nsta, nlta, position="centred", transform=transform,
                                         return_onsets=True)
    dev.free()
    env = PO.envelope(wf)
    sig = env if transform == "env" else env ** 2
    want = []
    for chans in rows:
        ons = np.array([O.centred_sta_lta(sig[c], nsta, nlta) for c in chans])
        want.append(np.clip(np.sqrt(np.sum(ons ** 2, axis=0) / len(ons)), 0.4, np.inf))
    np.testing.assert_allclose(got, np.stack(want), rtol=1e-8)
