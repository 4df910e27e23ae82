# -*- coding: utf-8 -*-
"""
Detect output stage on the device: ``qmb_detect_fetch_quantised`` against the
reference's host arithmetic -- ``ScanmSEED.append`` (io/scanmseed.py:102-130) applied to the
float64 detect result and ``Grid3D.index2coord(unravel=True)`` (lut/lut.py:142-243).  Integer
channels must be bit-exact; the files written from either path must be byte-identical.
"""
import datetime as dt
import json
import pathlib

import numpy as np
import pytest

from quakemigrate_b200 import scanmseed as sm
from quakemigrate_b200 import synthetic as syn
from quakemigrate_b200.lut import Grid3D, LambertConformalConic

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).resolve().parent / "golden"


def _grid(shape):
    meta = json.loads((GOLD / "golden_iceland_grid.json").read_text())
    proj = LambertConformalConic.from_proj_string(meta["grid_proj"])
    g = Grid3D(shape, meta["node_spacing"], meta["ll_corner"], grid2coord=proj.inverse)
    g.unit_conversion_factor = proj.unit_conversion_factor
    return g


def _detect(eng, grid=(13, 11, 9), n_samples=1500, n_rows=8, boost=False):
    cfg = syn.CONFIGS["cfg2"].scaled(grid=grid, n_samples=n_samples, n_rows=n_rows, name="out")
    prob = syn.make_problem(cfg)
    dev, onsets = eng.onsets_from_waveforms(prob.waveforms, prob.row_channels, prob.nsta, prob.nlta,
                                            position=cfg.position, return_onsets=True)
    if boost:  # a burst on every row that lifts exp(mean log onset) above the 21474 clip
        onsets = onsets.copy()
        onsets[:, prob.fsmp + 200:prob.fsmp + 600] = 1.0e5
        dev = eng.upload_onsets(onsets)
    lut = eng.upload_lut(prob.traveltimes)
    eng.detect_device(lut, dev, prob.fsmp, prob.lsmp, cfg.n_rows)
    return cfg, eng.detect_fetch(cfg.n_samples)


@pytest.mark.parametrize("boost", [False, True])
def test_quantised_fetch_matches_host_arithmetic(boost):
    from quakemigrate_b200.core import Engine

    with Engine(0) as eng:
        cfg, (max_coa, max_norm, idx) = _detect(eng, boost=boost)
        g = _grid(cfg.grid)
        zf = 1e3 * g.unit_conversion_factor
        q = eng.detect_fetch_quantised(cfg.n_samples, g.node_count, g.ll_corner, g.node_spacing, zf)
    if boost:
        assert max_coa.max() > 21474.0
    want_coa = np.round(np.minimum(max_coa, 21474.0) * 1e5).astype(np.int32)
    want_coan = np.round(np.minimum(max_norm, 21474.0) * 1e5).astype(np.int32)
    assert np.array_equal(q["coa"], want_coa)
    assert np.array_equal(q["coa_n"], want_coan)
    assert np.array_equal(q["ijk"], np.column_stack(np.unravel_index(idx, cfg.grid)))
    xyz = g.index2grid(idx, unravel=True)
    assert np.array_equal(q["xyz"], xyz)  # same two roundings as lut.py:177, no FMA
    assert np.array_equal(q["z"], np.round(xyz[:, 2] * zf).astype(np.int32))


def test_partial_outputs_and_errors():
    from quakemigrate_b200.core import Engine, QmbError

    with Engine(0) as eng:
        g = _grid((13, 11, 9))
        with pytest.raises(QmbError):  # nothing scanned yet
            eng.detect_fetch_quantised(10, g.node_count, g.ll_corner, g.node_spacing, 1e6)
        cfg, (max_coa, _, idx) = _detect(eng)
        q = eng.detect_fetch_quantised(cfg.n_samples, g.node_count, g.ll_corner, g.node_spacing, 1e6,
                                       want=("coa", "ijk"))
        assert set(q) == {"coa", "ijk"}
        assert np.array_equal(q["coa"], np.round(np.minimum(max_coa, 21474.0) * 1e5).astype(np.int32))
        with pytest.raises(QmbError):
            eng.detect_fetch_quantised(cfg.n_samples + 1, g.node_count, g.ll_corner, g.node_spacing, 1e6)


def test_files_from_device_and_host_paths_are_identical(tmp_path):
    from quakemigrate_b200.core import Engine

    t0 = dt.datetime(2014, 6, 29, 18, 42, 5)
    with Engine(0) as eng:
        cfg, (max_coa, max_norm, idx) = _detect(eng, n_samples=2500)
        g = _grid(cfg.grid)
        host = sm.ScanmSEED(tmp_path / "host", False, cfg.sampling_rate)
        host.append(t0, max_coa.copy(), max_norm.copy(), g.index2coord(idx, unravel=True),
                    g.unit_conversion_factor)
        dev = sm.ScanmSEED(tmp_path / "dev", False, cfg.sampling_rate)
        dev.append_from_engine(t0, eng, cfg.n_samples, g)
    a, b = host.write(), dev.write()
    assert a.name == b.name == "2014_180.scanmseed"
    assert a.read_bytes() == b.read_bytes()
    data, stats = sm.read_scanmseed(tmp_path / "dev", t0, t0 + dt.timedelta(seconds=24.99), 0.0,
                                    g.unit_conversion_factor)
    assert stats["npts"] == cfg.n_samples
    np.testing.assert_allclose(data["COA"].values, max_coa, atol=0.5e-5 * 1.0000001)


def test_continuous_compute_writes_the_same_file_either_way(tmp_path):
    """QuakeScan.continuous_compute (scan.py:406-467): three timesteps, the middle one without data."""
    from test_gpu_scan_api import _problem

    from quakemigrate_b200.core import Engine
    from quakemigrate_b200.onsets import STALTAOnset, WaveformData
    from quakemigrate_b200.scan import QuakeScan

    cfg, prob, lut, data = _problem()
    meta = json.loads((GOLD / "golden_iceland_grid.json").read_text())
    proj = LambertConformalConic.from_proj_string(meta["grid_proj"])
    lut.grid2coord, lut.unit_conversion_factor = proj.inverse, proj.unit_conversion_factor
    sr = cfg.sampling_rate
    with Engine(0) as eng:
        onset = STALTAOnset(engine=eng, sampling_rate=sr, position="classic")
        scan = QuakeScan(lut, onset, engine=eng, timestep=4.0)
        pre, post = scan.set_pads(scan.timestep)
        n_chunk = int(round((pre + post + scan.timestep) * sr))
        t_first = 1_404_067_325.0  # 2014-06-29 18:42:05 UTC
        chunks = []
        for i in range(3):
            if i == 1:
                chunks.append(None)
                continue
            o = i * int(scan.timestep * sr)
            ch = {st: {c: v[o:o + n_chunk] for c, v in comps.items()} for st, comps in data.channels.items()}
            chunks.append(WaveformData(ch, starttime=t_first + i * scan.timestep - pre, sampling_rate=sr))
        files = []
        for mode in (True, False):
            run = tmp_path / ("dev" if mode else "host")
            coalescence, availability = scan.continuous_compute(t_first, chunks, run, device_output=mode)
            assert coalescence.written and availability[1] is None and sum(availability[0].values()) > 0
            files.append(run / "detect" / "scanmseed" / "2014_180.scanmseed")
            av_file = run / "detect" / "availability" / "2014_180_StationAvailability.csv"
            lines = av_file.read_text().splitlines()
            assert len(lines) == 4 and lines[0].startswith(",ST000_P,") and lines[2].endswith(",0,0,0")
            assert lines[1].startswith("2014-06-29T18:42:05.000000Z,1,")
        scan.close()
    assert files[0].read_bytes() == files[1].read_bytes()
    df, stats = sm.read_scanmseed(tmp_path / "dev", t_first, t_first + 12.0, 0.0, lut.unit_conversion_factor)
    n = int(scan.timestep * sr)
    assert stats["npts"] == 3 * n
    assert (df["COA"].values[n:2 * n] == 0).all() and (df["COA"].values[:n] > 0).all()
