# -*- coding: utf-8 -*-
"""
The Icequake_Iceland example of the reference, as test infrastructure: rebuilds its travel-time
table (examples/Icequake_Iceland/iceland_lut.py: homogeneous model, create_lut.py:241-265) from the
committed facts, reads the raw event waveforms the reference's locate() run read, and parses the
.event files it wrote (tests/golden/iceland_locate, see tests/golden/make_golden_iceland_events.py).
"""
import csv
import datetime as dt
import json
import pathlib

import numpy as np

from quakemigrate_b200 import scanmseed as sm
from quakemigrate_b200.lut import LUT, LambertConformalConic

GOLD = pathlib.Path(__file__).resolve().parent / "golden" / "iceland_locate"


def facts() -> dict:
    return json.loads((GOLD / "lut_facts.json").read_text())


def stations():
    with open(GOLD / "iceland_stations.txt") as f:
        return list(csv.DictReader(f))


def build_lut() -> LUT:
    """compute_traveltimes(grid_spec, stations, method="homogeneous", vp, vs) without pyproj."""
    fx = facts()
    proj = LambertConformalConic.from_proj_string(fx["grid_proj"])
    spacing = np.asarray(fx["node_spacing"], dtype=float)
    ll = proj.forward([fx["ll_corner_lonlatdepth"]])[0]          # lut.py:88-89 (Grid3D.__init__)
    ur = proj.forward([fx["ur_corner_lonlatdepth"]])[0]
    node_count = (np.ceil((ur - ll) / spacing) + 1).astype(int)  # lut.py:92-94
    st = stations()
    # read_stations flips the sign of the elevation: depth is positive down (io/core.py:95)
    sxyz = proj.forward([[float(r["Longitude"]), float(r["Latitude"]), -float(r["Elevation"])] for r in st])
    grid = LUT({}, node_count, spacing, ll, phases=["P", "S"], grid2coord=proj.inverse)
    ijk = np.meshgrid(*[np.arange(n) for n in node_count], indexing="ij")
    g = grid.index2grid(np.column_stack([d.flatten() for d in ijk]))          # lut.py:329-336
    gx, gy, gz = [g[:, d].reshape(node_count) for d in range(3)]
    for i, r in enumerate(st):
        dx, dy, dz = gx - sxyz[i, 0], gy - sxyz[i, 1], gz - sxyz[i, 2]
        dist = np.sqrt(dx ** 2 + dy ** 2 + dz ** 2)                              # create_lut.py:262-263
        grid.traveltimes[r["Name"]] = {"P": dist / fx["vp"], "S": dist / fx["vs"]}
    grid.stations = [r["Name"] for r in st]
    grid.unit_conversion_factor = proj.unit_conversion_factor
    grid.coord2grid_fn = proj.forward
    return grid


def event_ids():
    return sorted(p.stem for p in GOLD.glob("*.event"))


def read_event_file(uid: str) -> dict:
    with open(GOLD / f"{uid}.event") as f:
        row = next(csv.DictReader(f))
    return row


def read_raw_waveforms(uid: str):
    """
    Returns (channels, start_us, sampling_rate): channels[station][component letter] = float64 array.
    All traces of the file share start, rate and length (checked).
    """
    traces = sm.read_mseed(GOLD / f"{uid}.m")
    start = {t.time_of(0) for t in traces}
    rate = {t.sampling_rate for t in traces}
    npts = {t.npts for t in traces}
    assert len(start) == len(rate) == len(npts) == 1, (start, rate, npts)
    channels = {}
    for t in traces:
        _, station, _, channel = t.id
        channels.setdefault(station, {})[channel[-1]] = np.asarray(t.data, dtype=np.double)
    return channels, start.pop(), rate.pop()


def utc_string(us: int) -> str:
    """obspy's str(UTCDateTime): 2014-06-29T18:42:08.388000Z."""
    d = dt.datetime(1970, 1, 1) + dt.timedelta(microseconds=int(us))
    return d.strftime("%Y-%m-%dT%H:%M:%S.%f") + "Z"


def locate_event(uid: str, engine, lut=None) -> dict:
    """
    The body of QuakeScan._locate_events for one event (scan.py:496-546) through quakemigrate_b200,
    starting from the raw traces the reference read; returns the columns of the .event file
    formatted like Event.write (io/event.py:441-505).
    """
    from quakemigrate_b200.onsets import STALTAOnset, WaveformData
    from quakemigrate_b200.scan import QuakeScan

    fx = facts()["locate"]
    lut = lut or build_lut()
    channels, start_us, raw_rate = read_raw_waveforms(uid)
    sr = fx["sampling_rate"]
    # util.resample (util.py:404-478), called per phase inside pre_process (stalta.py:189-191): every
    # trace is decimated 500 -> 250 Hz (detrend, taper, zero-phase low-pass, every second sample)
    from quakemigrate_b200 import preprocess as pp
    keys = [(s, c) for s in channels for c in channels[s]]
    block = np.stack([channels[s][c] for s, c in keys])
    if hasattr(engine, "preprocess"):
        dec = pp.decimate(block, raw_rate, sr, engine=engine)
    else:  # the oracle-backed stand-in of the CPU tests
        sos = pp.butter_sos("lowpass", float(sr) / 2.000001, raw_rate, 2)
        raw = engine.upload_waveforms(block)
        dec = engine.filter_waveforms(raw, sos, decimate=int(raw_rate / sr)).fetch()
    dec_channels = {}
    for (s, c), row in zip(keys, dec):
        dec_channels.setdefault(s, {})[c] = row
    data = WaveformData(dec_channels, starttime=0.0, sampling_rate=sr)
    data.stations = lut.stations  # the archive's station list: SKG09 has no data -> unavailable
    onset = STALTAOnset(engine=engine, position=fx["position"], sampling_rate=sr, phases=fx["phases"],
                        bandpass_filters=fx["bandpass_filters"], sta_lta_windows=fx["sta_lta_windows"])
    scan = QuakeScan(lut, onset, engine=engine, marginal_window=fx["marginal_window"], serve_on_device=False)
    scan.set_pads(4 * fx["marginal_window"])                                         # scan.py:494
    out = scan.compute_locate(data, fit_location=True)
    i = out["peak_sample"]
    loc = out["location"]
    otime_us = start_us + int(round(out["times"][i] * 1e6))
    xyz_dec = [6, 6, 4]  # io/event.py:484-495 with lut.precision = [.., .., 2] (25 m nodes) and km units
    cov_unc = loc["covariance"][1]
    res = {
        "DT": utc_string(otime_us), "COA": f"{out['max_coa'][i]:.4g}", "COA_NORM": f"{out['max_coa_n'][i]:.4g}",
        "COV_Err_XYZ": round(float(np.power(cov_unc[0] * cov_unc[1] * cov_unc[2], 1 / 3)), 4),   # event.py:260
    }
    for k, axis in enumerate("XYZ"):
        res[axis] = round(float(loc["spline"][k]), xyz_dec[k])
        res[f"GAU_{axis}"] = round(float(loc["gaussian"][0][k]), xyz_dec[k])
        res[f"GAU_Err{axis}"] = round(float(loc["gaussian"][1][k]), 4)
        res[f"COV_Err{axis}"] = round(float(cov_unc[k]), 4)
    res["_availability"] = dict(out["onset_data"].availability)
    res["_window"] = out["window"]
    scan.close()
    return res
