# -*- coding: utf-8 -*-
"""
Host logic of the detect loop without a GPU: QuakeScan._compute / continuous_compute,
STALTAOnset (with and without device pre-processing), the travel-time cache, the output files and
the error paths, driven through an oracle-backed stand-in for the engine
(tests/standin_engine.py -- test infrastructure, unreachable from the product).
Reference: quakemigrate/signal/scan.py:406-467, :594-647; signal/onsets/stalta.py:353-489.
"""
import datetime as dt

import numpy as np
import pytest

from oracle import preprocess_oracle as PO
from oracle import qm_oracle as O
from quakemigrate_b200 import scanmseed as sm
from quakemigrate_b200 import synthetic as syn
from quakemigrate_b200.lut import LUT
from quakemigrate_b200.onsets import STALTAOnset, WaveformData
from quakemigrate_b200.scan import LUTPhasesException, QuakeScan
from standin_engine import StandInEngine

SR = 50
T_FIRST = 1_404_067_325.0  # 2014-06-29 18:42:05 UTC


def _setup(n_sta=3, grid=(6, 5, 4), seconds=40.0, seed=0):
    rng = np.random.default_rng(seed)
    h = 0.5
    gx, gy, gz = np.meshgrid(*(np.arange(n) * h for n in grid), indexing="ij")
    sta = rng.random((n_sta, 2)) * np.array([grid[0], grid[1]]) * h
    tts = {}
    for k in range(n_sta):
        d = np.sqrt((gx - sta[k, 0]) ** 2 + (gy - sta[k, 1]) ** 2 + gz ** 2)
        tts[f"ST{k}"] = {"P": d / 5.0, "S": d / 2.9}
    lut = LUT(tts, grid, (h, h, h), (0.0, 0.0, 0.0))
    n = int(seconds * SR)
    channels = {f"ST{k}": {c: rng.standard_normal(n) * 50 + 1000.0 + 0.01 * np.arange(n) for c in "ZNE"}
                for k in range(n_sta)}
    return lut, channels


def _chunk(channels, o, n, start):
    return WaveformData({st: {c: v[o:o + n] for c, v in comps.items()} for st, comps in channels.items()},
                        starttime=start, sampling_rate=SR)


def test_compute_matches_the_reference_chain():
    lut, channels = _setup()
    eng = StandInEngine()
    onset = STALTAOnset(engine=eng, sampling_rate=SR, position="classic")
    scan = QuakeScan(lut, onset, engine=eng, timestep=10.0, serve_on_device=False)
    pre, post = scan.set_pads(scan.timestep)
    n = int(round((pre + post + scan.timestep) * SR))
    data = _chunk(channels, 0, n, T_FIRST - pre)
    time, max_coa, max_coa_n, coord, onset_data = scan._compute(data)
    assert time == T_FIRST and len(max_coa) == int(scan.timestep * SR)
    # scan.py:618-639 by hand
    rows, nsta, nlta = [], 11, 51
    for phase, comps in (("P", "Z"), ("S", "NE")):
        for st in channels:
            ons = np.array([O.overlapping_sta_lta(data.channels[st][c] ** 2, nsta, nlta) for c in comps])
            rows.append(np.clip(np.sqrt(np.sum(ons ** 2, axis=0) / len(ons)), 0.4, np.inf))
    tt = lut.serve_traveltimes(SR, onset_data.availability)
    fsmp, lsmp = syn.time2sample(pre, SR), syn.time2sample(post, SR)
    m4 = O.migrate(np.stack(rows), tt, fsmp, lsmp, 6, threads=2)
    w_max, w_norm, w_idx = O.find_max_coa(m4, threads=2)
    np.testing.assert_allclose(max_coa, w_max, rtol=1e-12)
    np.testing.assert_allclose(max_coa_n, w_norm, rtol=1e-12)
    np.testing.assert_allclose(coord, lut.index2coord(w_idx, unravel=True))
    assert list(onset_data.availability) == ["ST0_P", "ST1_P", "ST2_P", "ST0_S", "ST1_S", "ST2_S"]
    # the table is cached per availability pattern (the reference re-serves it every call)
    scan._compute(data)
    assert eng.uploads == 1
    del data.channels["ST1"]["N"], data.channels["ST1"]["E"]
    scan._compute(data)
    assert eng.uploads == 2
    scan.close()


def test_onset_with_bandpass_filters_runs_pre_process_per_phase():
    lut, channels = _setup(n_sta=2)
    eng = StandInEngine()
    filters = {"P": [2.0, 16.0, 2], "S": [1.0, 8.0, 3]}
    onset = STALTAOnset(engine=eng, sampling_rate=SR, position="centred", bandpass_filters=filters,
                        signal_transform="env_squared")
    data = _chunk(channels, 0, 1500, 0.0)
    onsets, od = onset.calculate_onsets(data)
    want = []
    for phase, comps in (("P", "Z"), ("S", "NE")):
        for st in channels:
            filt = [PO.pre_process(data.channels[st][c], SR, filters[phase]) for c in comps]
            ons = np.array([O.centred_sta_lta(PO.envelope(f) ** 2, 11, 51) for f in filt])
            want.append(np.clip(np.sqrt(np.sum(ons ** 2, axis=0) / len(ons)), 0.4, np.inf))
            np.testing.assert_allclose(od.filtered_waveforms.trace(st, comps[0], phase), filt[0], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(onsets, np.stack(want), rtol=1e-9)
    assert set(od.filtered_waveforms.to_host()) == {(st, c, ph) for ph, cs in (("P", "Z"), ("S", "NE"))
                                                    for st in channels for c in cs}


def test_continuous_compute_files(tmp_path):
    lut, channels = _setup(seconds=60.0)
    eng = StandInEngine()
    onset = STALTAOnset(engine=eng, sampling_rate=SR)
    scan = QuakeScan(lut, onset, engine=eng, timestep=8.0, serve_on_device=False)
    pre, post = scan.set_pads(scan.timestep)
    n = int(round((pre + post + scan.timestep) * SR))
    step = int(scan.timestep * SR)
    chunks = [_chunk(channels, 0, n, T_FIRST - pre), None, _chunk(channels, 2 * step, n, T_FIRST + 16.0 - pre)]
    del chunks[2].channels["ST2"]["Z"]
    outs = []
    for mode in (True, False):
        run = tmp_path / ("dev" if mode else "host")
        coalescence, availability = scan.continuous_compute(T_FIRST, chunks, run, device_output=mode)
        assert coalescence.written and availability[1] is None
        outs.append(((run / "detect" / "scanmseed" / "2014_180.scanmseed").read_bytes(),
                     (run / "detect" / "availability" / "2014_180_StationAvailability.csv").read_text()))
    assert outs[0] == outs[1]
    lines = outs[0][1].splitlines()
    assert lines[0] == ",ST0_P,ST1_P,ST2_P,ST0_S,ST1_S,ST2_S"
    assert lines[1] == "2014-06-29T18:42:05.000000Z,1,1,1,1,1,1"
    assert lines[2] == "2014-06-29T18:42:13.000000Z,0,0,0,0,0,0"
    assert lines[3] == "2014-06-29T18:42:21.000000Z,1,1,0,1,1,1"
    data, stats = sm.read_scanmseed(tmp_path / "dev", T_FIRST, T_FIRST + 24.0, 0.0, 1.0)
    assert stats["npts"] == 3 * step and stats["starttime"] == dt.datetime(2014, 6, 29, 18, 42, 5)
    assert (data["COA"].values[step:2 * step] == 0).all() and (data["COA"].values[:step] > 0).all()
    scan.close()


def test_error_paths():
    lut, channels = _setup()
    eng = StandInEngine()
    onset = STALTAOnset(engine=eng, sampling_rate=SR, phases=["P", "X"], channel_maps={"P": "Z", "X": "Z"},
                        sta_lta_windows={"P": [0.2, 1.0], "X": [0.2, 1.0]})
    scan = QuakeScan(lut, onset, engine=eng, serve_on_device=False)
    scan.set_pads(10.0)
    with pytest.raises(LUTPhasesException):  # scan.py:624-629
        scan._compute(_chunk(channels, 0, 1500, 0.0))
    onset = STALTAOnset(engine=eng, sampling_rate=SR)
    scan = QuakeScan(lut, onset, engine=eng, serve_on_device=False)
    pre, post = scan.set_pads(10.0)
    with pytest.raises(ValueError):  # lib.py:109-110: data shorter than the pads
        scan._compute(_chunk(channels, 0, int((pre + post) * SR) - 5, 0.0))
    with pytest.raises(TypeError):
        QuakeScan(lut, object())


def test_compute_locate_window_and_location(tmp_path):
    """compute_locate (io/event.py:421-439 window, scan.py:696-733 fits): the marginal window sits
    around the peak of max_coa, the map is the window's time-sum, and a planted source is found."""
    from quakemigrate_b200.scan import marginal_window_indices

    lut, channels = _setup(n_sta=4, grid=(9, 8, 6), seconds=40.0, seed=3)
    # plant an impulsive arrival from node (5, 3, 2) at t0 = 20 s on every channel
    src = (5, 3, 2)
    t0 = 20.0
    for st, comps in channels.items():
        for c, v in comps.items():
            tt = lut[st]["P" if c == "Z" else "S"][src]
            k = int(round((t0 + tt) * SR))
            v[k:k + 25] += 4000.0 * np.hanning(25) * np.sin(np.arange(25))
    eng = StandInEngine()
    onset = STALTAOnset(engine=eng, sampling_rate=SR, position="centred")
    scan = QuakeScan(lut, onset, engine=eng, marginal_window=1.0, serve_on_device=False)
    pre, post = scan.set_pads(20.0)
    o = int(round((10.0 - pre) * SR))
    n = int(round((pre + post + 20.0) * SR))
    data = _chunk(channels, o, n, 10.0 - pre)
    out = scan.compute_locate(data, return_map4d=True, fit_location=True)
    i0, i1 = out["window"]
    imax = out["peak_sample"]
    assert (i0, i1, imax) == marginal_window_indices(out["max_coa"], 1.0, SR)
    assert abs(out["times"][imax] - t0) < 0.2 and i1 - i0 == 2 * SR  # end-exclusive slice of 2 s + 1 sample
    assert out["map4d"].shape == tuple(lut.node_count) + (i1 - i0,)
    np.testing.assert_allclose(out["coa_map"], out["map4d"].sum(axis=-1), rtol=1e-12)
    peak = np.unravel_index(np.argmax(out["coa_map"]), out["coa_map"].shape)
    assert max(abs(a - b) for a, b in zip(peak, src)) <= 1
    loc = out["location"]
    want_xyz = lut.index2coord([list(src)])[0]
    for key in ("spline", "gaussian", "covariance"):
        xyz = loc[key] if key == "spline" else loc[key][0]
        assert np.abs(np.asarray(xyz) - want_xyz).max() < 1.0  # within two nodes (0.5 km spacing)
    scan.close()
