# -*- coding: utf-8 -*-
"""
Detect output stage: .scanmseed files.

The golden vector is the reference's OWN detect output,
``examples/benchmarks/Icequake_Iceland/2014_180.scanmseed`` (what
``tests/test_benchmarks.py:80-105`` of the reference compares a fresh detect() run against),
committed bit-identical as ``tests/golden/golden_2014_180.scanmseed`` together with the grid facts
of the example LUT (``tests/golden/make_golden_scanmseed.py``).  It pins

* the STEIM-2 codec and the miniSEED record framing byte for byte (read -> write == file),
* ``index2coord(unravel=True)`` + the Lambert projection + the x1e6 / x1e3*ucf quantisation of
  ``ScanmSEED.append`` (scanmseed.py:102-130): every X/Y/Z sample of the file sits on a node of
  the example grid, and re-deriving the three channels from those node indices gives the file's
  integers exactly.

The STEIM functions are host code inside libqmb200.so: they run without a GPU.
"""
import datetime as dt
import json
import pathlib

import numpy as np
import pytest

from quakemigrate_b200 import _lib, scanmseed as sm
from quakemigrate_b200.lut import Grid3D, LambertConformalConic

GOLD = pathlib.Path(__file__).resolve().parent / "golden"
FILE = GOLD / "golden_2014_180.scanmseed"


@pytest.fixture(scope="module")
def golden_traces():
    return sm.read_mseed(FILE)


@pytest.fixture(scope="module")
def grid():
    meta = json.loads((GOLD / "golden_iceland_grid.json").read_text())
    proj = LambertConformalConic.from_proj_string(meta["grid_proj"])
    g = Grid3D(meta["node_count"], meta["node_spacing"], meta["ll_corner"], grid2coord=proj.inverse)
    g.unit_conversion_factor = proj.unit_conversion_factor
    return g, proj, meta


def test_golden_file_reads(golden_traces):
    assert [t.station for t in golden_traces] == ["COA", "COA_N", "X", "Y", "Z"]
    for tr in golden_traces:
        assert tr.network == "NW" and tr.npts == 2500 and tr.sampling_rate == 250.0
        assert tr.starttime == dt.datetime(2014, 6, 29, 18, 42, 5)
        assert tr.endtime == dt.datetime(2014, 6, 29, 18, 42, 14, 996000)
    coa, coa_n = golden_traces[0].data / 1e5, golden_traces[1].data / 1e5
    assert 1.0 < coa.min() and coa.max() < 3.0  # geometric-mean onsets of noise + three icequakes
    assert (coa_n >= 1.0).all()                 # max * N / sum >= 1


def test_golden_file_roundtrip_bytes(golden_traces, tmp_path):
    """read -> write reproduces libmseed's records exactly (headers and STEIM-2 frames)."""
    out = tmp_path / "again.scanmseed"
    sm.write_mseed(out, golden_traces, encoding="STEIM2")
    assert out.read_bytes() == FILE.read_bytes()


def test_golden_record_split(golden_traces):
    """Samples per record follow from the packing alone: 1890+610, 1873+627, ..., 966+975+559."""
    blob = FILE.read_bytes()
    want = [r["n"] for r in sm._parse_records(blob)]
    got = [r["n"] for tr in golden_traces for r in sm._parse_records(sm.encode_trace(tr))]
    assert got == want == [1890, 610, 1873, 627, 1777, 723, 2169, 331, 966, 975, 559]


def test_golden_coordinates_sit_on_nodes(golden_traces, grid):
    g, proj, meta = grid
    x, y, z = (golden_traces[i].data for i in (2, 3, 4))
    ucf = g.unit_conversion_factor
    assert ucf == 1000.0  # +units=km
    coord = np.column_stack([x / 1e6, y / 1e6, z / (1e3 * ucf)])
    ijk_f = (proj.forward(coord) - g.ll_corner) / g.node_spacing
    ijk = np.rint(ijk_f).astype(int)
    # 1e-6 degree is 0.05-0.11 m, i.e. < 0.005 nodes at 25 m spacing
    assert np.abs(ijk_f - ijk).max() < 5e-3
    assert (ijk >= 0).all() and (ijk < np.array(meta["node_count"])).all()
    # and back: flat index -> index2coord(unravel=True) -> ScanmSEED._data2int == the file
    flat = np.ravel_multi_index(tuple(ijk.T), meta["node_count"])
    back = g.index2coord(flat, unravel=True)
    assert np.array_equal(sm.ScanmSEED._data2int(back[:, 0], 1e6), x)
    assert np.array_equal(sm.ScanmSEED._data2int(back[:, 1], 1e6), y)
    assert np.array_equal(sm.ScanmSEED._data2int(back[:, 2], 1e3 * ucf), z)


def test_projection_reproduces_grid_corner(grid):
    """The pickled ll_corner is pyproj's forward of grid_spec.ll_corner (iceland_lut.py:49)."""
    g, proj, meta = grid
    got = proj.forward(np.array([[-17.24, 64.322, -1.4]]))[0]
    np.testing.assert_allclose(got, meta["ll_corner"], atol=2e-9)  # km -> 2 micrometres
    got = proj.forward(np.array([[-17.204, 64.336, 0.0]]))[0]
    np.testing.assert_allclose(got, meta["ur_corner"], atol=2e-9)
    lonlat = proj.inverse(proj.forward(np.array([[-17.2, 64.33, -0.5]])))
    np.testing.assert_allclose(lonlat, [[-17.2, 64.33, -0.5]], atol=1e-12)


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("seed", range(4))
def test_steim_roundtrip_random(version, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 9000))
    scale = [3, 40, 300, 20000, 3_000_000, 200_000_000][int(rng.integers(0, 6))]
    steps = rng.integers(-scale, scale + 1, n)
    steps[rng.random(n) < 0.3] = 0
    if version == 1:
        steps[rng.random(n) < 0.01] = 2_000_000_000 // 4
    data = np.clip(np.cumsum(steps), -2**31 + 1, 2**31 - 1).astype(np.int32)
    if version == 2:  # keep every difference inside 30 bits
        d = np.clip(np.diff(data.astype(np.int64), prepend=data[0]), -(2**29), 2**29 - 1)
        data = (data[0] + np.cumsum(d) - d[0]).astype(np.int32)
    i, prev, out = 0, None, []
    while i < n:
        d0 = 0 if prev is None else int(data[i]) - prev
        body, cnt = sm.steim_encode(data[i:], d0, version, n_frames=int(rng.integers(1, 64)))
        assert cnt >= 1
        dec, got_d0 = sm.steim_decode(body, cnt, version)
        assert got_d0 == d0
        out.append(dec)
        prev, i = int(data[i + cnt - 1]), i + cnt
    assert np.array_equal(np.concatenate(out), data)


def test_steim2_field_boundaries():
    """Each field width at its limits, and the greedy group choice (7x4 ... 1x30)."""
    for lim, per_word in [(7, 7), (15, 6), (31, 5), (127, 4), (511, 3), (16383, 2), (2**29 - 1, 1)]:
        steps = np.array([0] + [lim, -lim - 1] * 21, dtype=np.int64)
        data = np.cumsum(steps).astype(np.int32)
        body, cnt = sm.steim_encode(data, 0, 2, n_frames=1)
        # frame 0 has 13 data words; the first word also carries d0 = 0
        assert cnt == min(len(data), 13 * per_word), (lim, cnt)
        dec, _ = sm.steim_decode(body, cnt, 2)
        assert np.array_equal(dec, data[:cnt])
    with pytest.raises(_lib.QmbError) as e:
        sm.steim_encode(np.array([0, 2**29], dtype=np.int32), 0, 2, 1)
    assert e.value.status == 3  # QMB_ERR_UNSUPPORTED -> the caller falls back to STEIM-1


def test_steim_decode_rejects_corruption(golden_traces):
    body, cnt = sm.steim_encode(golden_traces[0].data, 0, 2, 63)
    bad = bytearray(body)
    bad[200] ^= 0x10
    with pytest.raises(_lib.QmbError):
        sm.steim_decode(bytes(bad), cnt, 2)


def _synthetic_step(rng, n, grid_shape):
    coa = 1.0 + rng.random(n) * 3.0
    coa_n = coa * (1.0 + rng.random(n))
    idx = rng.integers(0, int(np.prod(grid_shape)), n)
    return coa, coa_n, idx


def test_scanmseed_append_write_read(tmp_path, grid):
    """detect(): four timesteps appended, one day file, read back == round(x * factor) / factor."""
    g, proj, meta = grid
    rng = np.random.default_rng(0)
    sr, n = 50, 3000
    scan = sm.ScanmSEED(tmp_path, continuous_write=False, sampling_rate=sr)
    t0 = dt.datetime(2014, 6, 29, 18, 0, 0)
    chunks = []
    for i in range(4):
        coa, coa_n, idx = _synthetic_step(rng, n, meta["node_count"])
        if i == 2:
            coa[5] = 99999.0  # clipped at 21474 (scanmseed.py:103)
        coord = g.index2coord(idx, unravel=True)
        scan.append(t0 + dt.timedelta(seconds=i * n / sr), coa, coa_n, coord, g.unit_conversion_factor)
        assert coa.max() <= 21474.0  # in-place clip like the reference
        chunks.append((coa.copy(), coa_n.copy(), coord))
    assert len(scan.stream) == 5 and scan.stream[0].npts == 4 * n and not scan.written
    file = scan.write()
    assert file.name == "2014_180.scanmseed"
    data, stats = sm.read_scanmseed(tmp_path, t0, t0 + dt.timedelta(seconds=4 * n / sr), 0.0,
                                    g.unit_conversion_factor)
    assert stats["npts"] == 4 * n and stats["station"] == "COA" and stats["sampling_rate"] == sr
    coa = np.concatenate([c[0] for c in chunks])
    coord = np.concatenate([c[2] for c in chunks])
    np.testing.assert_array_equal(data["COA"].values, np.round(coa * 1e5) / 1e5)
    np.testing.assert_array_equal(data["X"].values, np.round(coord[:, 0] * 1e6) / 1e6)
    np.testing.assert_array_equal(data["Z"].values, np.round(coord[:, 2] * 1e6) / 1e6)
    assert data["COA"].max() == 21474.0
    assert data["DT"].iloc[0] == t0 and data["DT"].iloc[1] - data["DT"].iloc[0] == dt.timedelta(seconds=1 / sr)
    # trimmed read with a pad
    sub, st2 = sm.read_scanmseed(tmp_path, t0 + dt.timedelta(seconds=60), t0 + dt.timedelta(seconds=90),
                                 5.0, g.unit_conversion_factor)
    assert st2["npts"] == 40 * sr + 1 and sub["DT"].iloc[0] == t0 + dt.timedelta(seconds=55)
    np.testing.assert_array_equal(sub["COA"].values, data["COA"].values[55 * sr:95 * sr + 1])


def test_scanmseed_day_boundary_and_empty(tmp_path):
    """scanmseed.py:131-147: a stream ending at midnight - delta is written and reset; one that
    crosses midnight is split; empty() timesteps are zeros."""
    sr, step = 2, 3600.0
    scan = sm.ScanmSEED(tmp_path, continuous_write=False, sampling_rate=sr)
    t0 = dt.datetime(2020, 12, 31, 22, 0, 0)
    n = int(step * sr)
    ones = lambda v: np.full(n, float(v))  # noqa: E731
    coord = np.zeros((n, 3))
    scan.append(t0, ones(1), ones(1), coord, 1.0)
    assert not scan.written
    scan.append(t0 + dt.timedelta(hours=1), ones(2), ones(2), coord, 1.0)
    assert scan.written and scan.stream == []  # ended exactly at midnight - delta
    assert (tmp_path / "detect" / "scanmseed" / "2020_366.scanmseed").exists()
    # next day: half an hour late start so that the third step straddles midnight
    t1 = dt.datetime(2021, 1, 1, 22, 30, 0)
    scan.append(t1, ones(3), ones(3), coord, 1.0)
    scan.empty(t1, step, 1, "no data", 1.0)  # 23:30 .. 00:30 of zeros
    assert (tmp_path / "detect" / "scanmseed" / "2021_001.scanmseed").exists()
    assert scan.stream[0].starttime == dt.datetime(2021, 1, 2) and scan.stream[0].npts == n // 2
    assert not scan.written
    scan.write()
    day1, _ = sm.read_scanmseed(tmp_path, t1, dt.datetime(2021, 1, 2, 0, 29, 59, 500000), 0.0, 1.0)
    assert len(day1) == 2 * n
    assert (day1["COA"].values[:n] == 3.0).all() and (day1["COA"].values[n:] == 0.0).all()
    with pytest.raises(sm.NoScanMseedDataException):
        sm.read_scanmseed(tmp_path, dt.datetime(2019, 1, 1), dt.datetime(2019, 1, 1, 1), 0.0, 1.0)


def test_scanmseed_continuous_write_and_steim1_fallback(tmp_path):
    scan = sm.ScanmSEED(tmp_path, continuous_write=True, sampling_rate=10)
    t0 = dt.datetime(2022, 3, 1, 0, 0, 0)
    n = 100
    coa = np.ones(n)
    coord = np.zeros((n, 3))
    coord[::2, 2], coord[1::2, 2] = 300.0, -300.0  # x 1e3 * ucf: differences of 6e8 > 2**29
    scan.append(t0, coa, coa.copy(), coord, 1000.0)
    assert scan.written
    file = tmp_path / "detect" / "scanmseed" / "2022_060.scanmseed"
    recs = sm._parse_records(file.read_bytes())
    assert {r["version"] for r in recs} == {1}  # the whole file fell back to STEIM-1 (:219-224)
    data, _ = sm.read_scanmseed(tmp_path, t0, t0 + dt.timedelta(seconds=9.9), 0.0, 1000.0)
    assert np.array_equal(data["Z"].values, coord[:, 2])


def test_availability_file_roundtrip_bytes(tmp_path):
    """read -> write reproduces the reference's golden StationAvailability.csv byte for byte
    (pandas `to_csv` of a frame indexed by UTCDateTime, io/availability.py:123-152)."""
    import shutil

    from quakemigrate_b200 import availability as av

    gold = GOLD / "golden_2014_180_StationAvailability.csv"
    src = tmp_path / "a" / "detect" / "availability"
    src.mkdir(parents=True)
    shutil.copyfile(gold, src / "2014_180_StationAvailability.csv")
    t0 = dt.datetime(2014, 6, 29, 18, 42, 5)
    index, columns, table = av.read_availability(tmp_path / "a", t0, t0 + dt.timedelta(seconds=10))
    assert index == [t0 + dt.timedelta(seconds=2.5 * i) for i in range(4)]  # timestep 2.5 s (iceland_detect.py)
    assert columns[:2] == ["SKR01_P", "SKR02_P"] and columns[-1] == "SKG13_S" and table.shape == (4, 26)
    assert table.sum(axis=1).tolist() == [24, 24, 24, 24]  # SKG09 is down: 24 of 26 rows (cfg 1)
    rows = [dict(zip(columns, r)) for r in table]
    files = av.write_availability(tmp_path / "b", index, rows)
    assert [f.name for f in files] == ["2014_180_StationAvailability.csv"]
    assert files[0].read_bytes() == gold.read_bytes()
    # steps without data are rows of zeros; files split by day
    times = [dt.datetime(2020, 12, 31, 23, 59, 0), dt.datetime(2021, 1, 1, 0, 1, 0)]
    files = av.write_availability(tmp_path / "c", times, [None, rows[0]], columns=columns)
    assert sorted(f.name for f in files) == ["2020_366_StationAvailability.csv", "2021_001_StationAvailability.csv"]
    _, _, tab = av.read_availability(tmp_path / "c", times[0], times[1])
    assert tab[0].sum() == 0 and tab[1].sum() == 24
    with pytest.raises(av.NoStationAvailabilityDataException):
        av.read_availability(tmp_path / "c", dt.datetime(2019, 1, 1), dt.datetime(2019, 1, 2))


def test_mseed_gaps_and_trimmed_reads(tmp_path):
    """Records of one channel that do not follow each other become separate traces (obspy's
    read + merge(method=-1) keeps gaps); starttime / endtime trim to the samples inside."""
    sr = 128  # 7.8125 ms: record start times are not multiples of 0.1 ms (BTIME truncation)
    rng = np.random.default_rng(4)
    a = rng.integers(-50000, 50000, 5000).astype(np.int32)
    b = rng.integers(-5, 5, 3000).astype(np.int32)
    t0 = dt.datetime(2021, 3, 4, 5, 6, 7)
    t1 = t0 + dt.timedelta(seconds=len(a) / sr + 10.0)  # a 10 s gap
    file = tmp_path / "gappy.mseed"
    sm.write_mseed(file, [sm.Trace(a, "COA", t0, sr), sm.Trace(b, "COA", t1, sr)])
    got = sm.read_mseed(file)
    assert len(got) == 2 and got[0].npts == 5000 and got[1].npts == 3000
    assert np.array_equal(got[0].data, a) and np.array_equal(got[1].data, b)
    assert got[0].starttime == t0 and abs((got[1].starttime - t1).total_seconds()) < 1e-4
    # several records per trace were needed and their (truncated) start times still chain up
    assert len(sm._parse_records(file.read_bytes())) > 2
    # trimmed read: 2 s .. 4 s of the first trace, end inclusive
    sub = sm.read_mseed(file, starttime=t0 + dt.timedelta(seconds=2), endtime=t0 + dt.timedelta(seconds=4))
    assert len(sub) == 1 and sub[0].npts == 2 * sr + 1
    assert np.array_equal(sub[0].data, a[2 * sr:4 * sr + 1])
    # nothing inside the gap
    assert sm.read_mseed(file, starttime=t0 + dt.timedelta(seconds=42), endtime=t0 + dt.timedelta(seconds=45)) == []


def test_time_inputs():
    us = sm.to_us(dt.datetime(2014, 6, 29, 18, 42, 5))
    assert us == 1_404_067_325_000_000
    assert sm.to_us(1_404_067_325.0) == us and sm.to_us(np.datetime64("2014-06-29T18:42:05")) == us
    aware = dt.datetime(2014, 6, 29, 20, 42, 5, tzinfo=dt.timezone(dt.timedelta(hours=2)))
    assert sm.to_us(aware) == us and sm.from_us(us) == dt.datetime(2014, 6, 29, 18, 42, 5)
    with pytest.raises(TypeError):
        sm.to_us("2014-06-29")
    with pytest.raises(ValueError):
        sm.write_mseed("/dev/null", [sm.Trace(np.zeros(4, np.int32), "COA", 0, 12.5)])
