# -*- coding: utf-8 -*-
"""
Plugin-level parity: quakemigrate_b200.scan.QuakeScan._compute (GPU) against a line-by-line CPU
restatement of the reference's QuakeScan._compute (scan.py:618-647) built from the oracle.
"""

import numpy as np
import pytest

from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _problem():
    cfg = syn.CONFIGS["cfg2"].scaled(grid=(14, 12, 9), n_samples=1200, n_rows=8, name="api")
    prob = syn.make_problem(cfg)
    n_sta = len(prob.stations_xyz)
    h = cfg.node_spacing_km
    from quakemigrate_b200.lut import LUT
    from quakemigrate_b200.onsets import WaveformData
    nx, ny, nz = cfg.grid
    gx, gy, gz = np.meshgrid(np.arange(nx) * h, np.arange(ny) * h, np.arange(nz) * h, indexing="ij")
    tts = {}
    for k in range(n_sta):
        d = np.sqrt((gx - prob.stations_xyz[k, 0]) ** 2 + (gy - prob.stations_xyz[k, 1]) ** 2 + gz ** 2)
        tts[f"ST{k:03d}"] = {"P": d / cfg.vp, "S": d / cfg.vs}
    lut = LUT(tts, cfg.grid, (h, h, h), (0.0, 0.0, 0.0))
    channels = {f"ST{k:03d}": {"Z": prob.waveforms[3 * k], "N": prob.waveforms[3 * k + 1],
                               "E": prob.waveforms[3 * k + 2]} for k in range(n_sta)}
    # one station lost its horizontals: exercises availability + LUT re-serving
    del channels["ST002"]["N"], channels["ST002"]["E"]
    data = WaveformData(channels, starttime=1000.0, sampling_rate=cfg.sampling_rate)
    return cfg, prob, lut, data


def _reference_compute(cfg, lut, data, pre_pad, post_pad, position):
    """scan.py:618-639 with the reference's own building blocks (oracle back-ends)."""
    sr = cfg.sampling_rate
    nsta, nlta = syn.time2sample(cfg.sta, sr) + 1, syn.time2sample(cfg.lta, sr) + 1
    fn = O.overlapping_sta_lta if position == "classic" else O.centred_sta_lta
    onsets, availability = [], {}
    for phase, comps in (("P", "Z"), ("S", "NE")):
        for st in data.stations:
            have = [c for c in comps if c in data.channels[st]]
            availability[f"{st}_{phase}"] = 1 if have else 0
            if not have:
                continue
            ons = np.array([fn(data.channels[st][c] ** 2, nsta, nlta) for c in have])
            onset = np.sqrt(np.sum([o ** 2 for o in ons], axis=0) / len(ons))
            onsets.append(np.clip(onset, 0.4, np.inf))
    onsets = np.stack(onsets, axis=0)
    traveltimes = lut.serve_traveltimes(sr, availability)
    fsmp, lsmp = syn.time2sample(pre_pad, sr), syn.time2sample(post_pad, sr)
    avail = int(np.sum(list(availability.values())))
    map4d = O.migrate(onsets, traveltimes, fsmp, lsmp, avail, threads=4)
    max_coa, max_coa_n, max_idx = O.find_max_coa(map4d, threads=4)
    return onsets, availability, map4d, max_coa, max_coa_n, max_idx, lut.index2coord(max_idx, unravel=True)


@pytest.mark.parametrize("position", ["classic", "centred"])
def test_compute_detect_and_locate(position):
    from quakemigrate_b200.core import Engine
    from quakemigrate_b200.onsets import STALTAOnset
    from quakemigrate_b200.scan import QuakeScan, marginal_window_indices

    cfg, prob, lut, data = _problem()
    eng = Engine(0)
    onset = STALTAOnset(engine=eng, sampling_rate=cfg.sampling_rate, position=position,
                        sta_lta_windows={"P": [cfg.sta, cfg.lta], "S": [cfg.sta, cfg.lta]})
    scan = QuakeScan(lut, onset, engine=eng, marginal_window=1.0)
    # pads exactly as the synthetic problem was built (onset.pad on the scan timespan)
    pre_pad, post_pad = scan.set_pads(cfg.n_samples / cfg.sampling_rate)
    assert (syn.time2sample(pre_pad, 100), syn.time2sample(post_pad, 100)) == (prob.fsmp, prob.lsmp)

    ref = _reference_compute(cfg, lut, data, pre_pad, post_pad, position)
    r_onsets, r_avail, r_map4d, r_max, r_norm, r_idx, r_coord = ref

    time, max_coa, max_coa_n, coord, onset_data = scan._compute(data)
    assert time == data.starttime + pre_pad
    assert list(onset_data.availability.items()) == list(r_avail.items())
    assert np.array_equal(np.argmax(r_map4d.reshape(-1, r_map4d.shape[-1]), axis=0), r_idx)
    np.testing.assert_allclose(max_coa, r_max, rtol=1e-8)
    np.testing.assert_allclose(max_coa_n, r_norm, rtol=1e-5)
    np.testing.assert_allclose(coord, r_coord, rtol=0, atol=0)      # identical arg-max -> identical coords

    # reference-shaped locate: the whole volume
    times, mc, mcn, crd, map4d, _ = scan._compute(data, event=object())
    assert map4d.shape == r_map4d.shape
    np.testing.assert_allclose(map4d, r_map4d, rtol=1e-5)

    # fast locate: marginal window only (event.py:421-439, scan.py:720-721)
    out = scan.compute_locate(data, return_map4d=True, fit_location=True)
    i0, i1, imax = marginal_window_indices(r_max, 1.0, cfg.sampling_rate)
    assert out["window"] == (i0, i1) and out["peak_sample"] == imax
    want = r_map4d[..., i0:i1]
    np.testing.assert_allclose(out["map4d"], want, rtol=1e-5)
    coa_map = out["coa_map"] / np.nanmax(out["coa_map"])
    r_coa_map = want.sum(axis=-1)
    np.testing.assert_allclose(coa_map, r_coa_map / np.nanmax(r_coa_map), rtol=1e-5)
    # location estimates (scan.py:696-733) against the oracle restatement on the reference map
    from quakemigrate_b200 import locate
    r_norm_map = r_coa_map / np.nanmax(r_coa_map)
    e, cov = O.np_covfit3d(r_norm_map, lut.node_spacing)
    np.testing.assert_allclose(out["location"]["covariance"][0], lut.ll_corner + e, rtol=1e-6)
    np.testing.assert_allclose(out["location"]["covariance"][1], np.diag(np.sqrt(np.abs(cov))), rtol=1e-4)
    r_sm = O.np_gaufilt3d(r_norm_map)
    r_ijk = np.unravel_index(np.nanargmax(r_sm), r_sm.shape)
    g_idx, g_sig = locate.gaussian_fit(r_sm, r_ijk, np.nanmean(r_sm))
    np.testing.assert_allclose(out["location"]["gaussian"][0], lut.index2coord([list(g_idx)])[0], rtol=1e-5)
    np.testing.assert_allclose(out["location"]["gaussian"][1], g_sig * lut.node_spacing, rtol=1e-3)
    scan.close()
    eng.close()


def test_device_served_table_equals_host_served():
    """qmb_lut_serve == qmb_lut_upload(LUT.serve_traveltimes(...)) (lut.py:502-538), incl. x-slabs."""
    from quakemigrate_b200.core import Engine
    cfg, prob, lut, data = _problem()
    eng = Engine(0)
    avail = {f"ST{k:03d}_{ph}": (0 if (k, ph) in ((1, "P"), (3, "S")) else 1) for ph in ("P", "S") for k in range(4)}
    host = lut.serve_traveltimes(cfg.sampling_rate, avail)
    grids, index = lut.to_device(eng)
    rows = lut.rows_for(avail, index)
    a = eng.upload_lut(host)
    b = eng.serve_lut(grids, rows, cfg.sampling_rate)
    ia, ib = a.info(), b.info()
    for k in ("n_rows", "max_tau", "max_tile_span", "n_tiles", "n_nodes_local"):
        assert ia[k] == ib[k], k
    onsets, _ = syn.random_problem(9, cfg.grid, host.shape[-1], 300, 10, int(host.max()) + 5)
    ra = eng.detect(a, onsets, 10, int(host.max()) + 5, host.shape[-1])
    rb = eng.detect(b, onsets, 10, int(host.max()) + 5, host.shape[-1])
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    # slab form: the second half of the x-planes
    x0 = cfg.grid[0] // 2
    c = eng.serve_lut(grids, rows, cfg.sampling_rate, x_range=(x0, cfg.grid[0]))
    d = eng.upload_lut(host[x0:], node_offset=x0 * cfg.grid[1] * cfg.grid[2], n_nodes_global=cfg.n_nodes)
    rc = eng.detect(c, onsets, 10, int(host.max()) + 5, host.shape[-1])
    rd = eng.detect(d, onsets, 10, int(host.max()) + 5, host.shape[-1])
    for x, y in zip(rc, rd):
        assert np.array_equal(x, y)
    with pytest.raises(KeyError):
        lut.rows_for({"ST000_X": 1}, index)
    eng.close()


def test_onset_timespan_nulls_taper_windows():
    """STALTAOnset.calculate_onsets(data, timespan): _trim_taper_pad semantics (stalta.py:550-584)."""
    from quakemigrate_b200.core import Engine
    from quakemigrate_b200.onsets import STALTAOnset
    cfg, prob, lut, data = _problem()
    eng = Engine(0)
    onset = STALTAOnset(engine=eng, sampling_rate=cfg.sampling_rate,
                        sta_lta_windows={"P": [cfg.sta, cfg.lta], "S": [cfg.sta, cfg.lta]})
    onset.post_pad = lut.max_traveltime
    plain, od = onset.calculate_onsets(data)
    nulled, od2 = onset.calculate_onsets(data, timespan=8.0)
    assert od.device_onsets is not None and od2.device_onsets is None
    pre_pad, _ = onset.pad(8.0)
    taper = syn.time2sample(pre_pad - onset.pre_pad, cfg.sampling_rate)
    nsta, nlta = prob.nsta, prob.nlta
    want = plain.copy()
    want[:, : taper + nlta - 1] = 1.0
    want[:, -(nsta + taper):] = 1.0
    assert np.array_equal(nulled, want)
    eng.close()


def test_decimate_and_upsample_on_the_device():
    """LUT.decimate (lut.py:102-140) and util.upsample's interpolation (util.py:544-551) on resident
    data: bit-identical to the host forms."""
    from quakemigrate_b200.core import Engine
    from quakemigrate_b200.lut import LUT
    from quakemigrate_b200.preprocess import upsample
    rng = np.random.default_rng(4)
    shape = (17, 14, 11)
    tts = {f"S{k}": {"P": rng.random(shape), "S": rng.random(shape) * 1.7} for k in range(3)}
    lut = LUT(tts, shape, [0.5, 0.5, 0.5], [0.0, 0.0, -1.0])
    with Engine(0) as eng:
        grids, index = lut.to_device(eng)
        for df in ([2, 3, 4], [1, 1, 1], [5, 5, 5]):
            dec = lut.decimate(df)
            dgr = grids.decimate(df)
            assert dgr.shape == (6,) + tuple(dec.node_count)
            assert list(lut.decimated_grid(df).node_count) == list(dec.node_count)
            avail = {f"{st}_{ph}": 1 for ph in ("P", "S") for st in tts}
            rows = LUT.rows_for(avail, index)
            dl = eng.serve_lut(dgr, rows, 100.0)
            want = dec.serve_traveltimes(100.0, avail)
            onsets = np.clip(rng.gamma(2.0, 0.6, size=(6, 700)), 0.4, None)
            ref = eng.upload_lut(want)
            a = eng.detect(dl, onsets, 10, 200, 6)
            b = eng.detect(ref, onsets, 10, 200, 6)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
            dgr.free()
        grids.free()
        wf = rng.standard_normal((5, 333))
        dev = eng.upload_waveforms(wf)
        for up in (1, 2, 5):
            got = dev.upsample(up)
            want = np.stack([upsample(row, up) for row in wf])
            assert np.array_equal(got.fetch(), want)
            got.free()
