# -*- coding: utf-8 -*-
"""
The detect OUTPUT stage: what ``QuakeScan.detect`` does with the ``_compute`` result
(``quakemigrate/signal/scan.py:443-456``), mirrored from ``quakemigrate/io/scanmseed.py``:

* :class:`ScanmSEED` -- ``append`` (clip at 21474, scale by 1e5 / 1e6 / 1e3*ucf, round, int32;
  scanmseed.py:74-150), ``empty`` (:152-185), ``write`` (:187-224, day files
  ``<year>_<julday>.scanmseed``), plus :meth:`ScanmSEED.append_from_engine`, which takes the
  integer channels straight from the device (``qmb_detect_fetch_quantised``) so the float64
  traces never cross PCIe;
* :func:`read_scanmseed` (:243-327) -- back to a ``pandas.DataFrame`` with columns
  ``DT, COA, COA_N, X, Y, Z``.

The reference stores these files as miniSEED with STEIM-2 compression through obspy/libmseed;
neither is available offline, so the record framing (48-byte fixed header + blockette 1000, 4096-byte
records, 63 frames) is restated here and the STEIM packing lives in ``libqmb200.so``
(``qmb_steim_encode`` / ``qmb_steim_decode``, host code).  Both are pinned byte-for-byte by the
reference's own golden file ``examples/benchmarks/Icequake_Iceland/2014_180.scanmseed`` (copied
bit-identical into ``tests/golden/``; see ``tests/test_scanmseed.py``): reading it and writing it
back reproduces the file exactly.

Times are ``datetime.datetime`` (UTC, naive) at the interface and integer microseconds inside.
"""

from __future__ import annotations

import ctypes
import datetime as _dt
import pathlib
import struct

import numpy as np

from . import _lib

__all__ = ["ScanmSEED", "read_scanmseed", "read_mseed", "write_mseed", "Trace", "steim_encode",
           "steim_decode", "NoScanMseedDataException"]

_EPOCH = _dt.datetime(1970, 1, 1)
_ENCODINGS = {"STEIM1": (10, 1), "STEIM2": (11, 2)}
_CHANNELS = ("COA", "COA_N", "X", "Y", "Z")
RECORD_LENGTH = 4096
DATA_OFFSET = 64


class NoScanMseedDataException(Exception):
    """No .scanmseed data found for the requested period (``util.NoScanMseedDataException``)."""

    def __init__(self):
        super().__init__("QuakeMigrate Error: No .scanmseed data found.")


# ------------------------------------------------------------------------------------------------
# time helpers (integer microseconds since 1970-01-01 UTC)
# ------------------------------------------------------------------------------------------------
def to_us(t) -> int:
    """datetime / numpy datetime64 / float seconds / int microseconds -> int microseconds."""
    if isinstance(t, (int, np.integer)):
        return int(t)
    if isinstance(t, (float, np.floating)):
        return int(round(float(t) * 1e6))
    if isinstance(t, np.datetime64):
        return int(t.astype("datetime64[us]").astype(np.int64))
    if isinstance(t, _dt.datetime):
        if t.tzinfo is not None:
            t = t.astimezone(_dt.timezone.utc).replace(tzinfo=None)
        d = t - _EPOCH
        return (d.days * 86400 + d.seconds) * 1_000_000 + d.microseconds
    if hasattr(t, "timestamp"):  # obspy.UTCDateTime
        return int(round(float(t.timestamp) * 1e6))
    raise TypeError(f"cannot interpret {type(t).__name__} as a time")


def from_us(us: int) -> _dt.datetime:
    return _EPOCH + _dt.timedelta(microseconds=int(us))


def _day_start_us(us: int) -> int:
    return (us // 86_400_000_000) * 86_400_000_000


def _julday(us: int):
    d = from_us(us)
    return d.year, d.timetuple().tm_yday


# ------------------------------------------------------------------------------------------------
# STEIM through the library
# ------------------------------------------------------------------------------------------------
def steim_encode(samples: np.ndarray, d0: int = 0, version: int = 2, n_frames: int = 63):
    """Pack as many of ``samples`` (int32) as fit ``n_frames`` frames: (bytes, n_packed)."""
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    out = ctypes.create_string_buffer(n_frames * 64)
    n_packed = ctypes.c_int64(0)
    _lib.check(_lib.load().qmb_steim_encode(samples, samples.size, int(d0), int(version),
                                            int(n_frames), ctypes.cast(out, ctypes.c_void_p),
                                            ctypes.byref(n_packed)))
    return out.raw, n_packed.value


def steim_decode(frames: bytes, n_samples: int, version: int = 2):
    """Decode one record's data section: (int32 samples, first difference)."""
    out = np.empty(n_samples, dtype=np.int32)
    d0 = ctypes.c_int32(0)
    buf = ctypes.create_string_buffer(bytes(frames), len(frames))
    _lib.check(_lib.load().qmb_steim_decode(ctypes.cast(buf, ctypes.c_void_p), len(frames) // 64,
                                            int(version), int(n_samples), out, ctypes.byref(d0)))
    return out, d0.value


# ------------------------------------------------------------------------------------------------
# miniSEED 2 records
# ------------------------------------------------------------------------------------------------
class Trace:
    """A contiguous int32 series with the few header fields .scanmseed files use."""

    def __init__(self, data, station, starttime, sampling_rate, network="NW", location="",
                 channel=""):
        self._chunks = [np.ascontiguousarray(data, dtype=np.int32)]
        self._npts = int(self._chunks[0].size)
        self.station, self.network, self.location, self.channel = station, network, location, channel
        self.start_us = to_us(starttime)
        self.sampling_rate = float(sampling_rate)

    @property
    def data(self) -> np.ndarray:
        """The samples; pieces appended with :meth:`extend` are joined on first use."""
        if len(self._chunks) > 1:
            self._chunks = [np.concatenate(self._chunks)]
        return self._chunks[0]

    @data.setter
    def data(self, value):
        self._chunks = [np.ascontiguousarray(value, dtype=np.int32)]
        self._npts = int(self._chunks[0].size)

    def extend(self, more: np.ndarray) -> None:
        """Append samples without copying what is already there (a day of detect() appends
        hundreds of timesteps; records of a file arrive one by one)."""
        more = np.ascontiguousarray(more, dtype=np.int32)
        self._chunks.append(more)
        self._npts += int(more.size)

    @property
    def id(self):
        return (self.network, self.station, self.location, self.channel)

    @property
    def npts(self):
        return self._npts

    @property
    def delta_us(self):
        return 1e6 / self.sampling_rate

    def time_of(self, i: int) -> int:
        """Microsecond timestamp of sample ``i``."""
        return self.start_us + int(round(i * 1e6 / self.sampling_rate))

    @property
    def end_us(self):
        return self.time_of(self.npts - 1)

    @property
    def starttime(self):
        return from_us(self.start_us)

    @property
    def endtime(self):
        return from_us(self.end_us)

    def times_us(self) -> np.ndarray:
        return self.start_us + np.rint(np.arange(self.npts) * (1e6 / self.sampling_rate)).astype(np.int64)

    def slice(self, start_us=None, end_us=None) -> "Trace":
        """Samples with start <= t <= end (obspy ``Trace.slice``, nearest sample inside)."""
        i0, i1 = 0, self.npts
        tol = 1  # microsecond
        if start_us is not None:
            i0 = max(0, int(np.ceil((start_us - tol - self.start_us) * self.sampling_rate / 1e6)))
        if end_us is not None:
            i1 = min(self.npts, int(np.floor((end_us + tol - self.start_us) * self.sampling_rate / 1e6)) + 1)
        i1 = max(i0, i1)
        return Trace(self.data[i0:i1], self.station, self.time_of(i0), self.sampling_rate,
                     self.network, self.location, self.channel)


def _rate_factors(sampling_rate: float):
    sr = float(sampling_rate)
    if sr == int(sr) and 0 < sr <= 32767:
        return int(sr), 1
    if sr > 0 and 1.0 / sr == int(1.0 / sr) and 1.0 / sr <= 32767:
        return -int(1.0 / sr), 1
    raise ValueError(f"sampling rate {sampling_rate} Hz is not representable here (integer Hz or "
                     "integer-second periods only)")


def _rate_from_factors(fact: int, mult: int) -> float:
    if fact == 0 or mult == 0:
        return 0.0
    if fact > 0 and mult > 0:
        return float(fact * mult)
    if fact > 0 and mult < 0:
        return -float(fact) / mult
    if fact < 0 and mult > 0:
        return -float(mult) / fact
    return 1.0 / (fact * mult)


def _pack_header(seq, tr: Trace, start_us, n_samples, encoding_code):
    t = from_us(start_us)
    fact, mult = _rate_factors(tr.sampling_rate)
    head = f"{seq:06d}".encode() + b"D "
    head += tr.station.ljust(5)[:5].encode() + tr.location.ljust(2)[:2].encode()
    head += tr.channel.ljust(3)[:3].encode() + tr.network.ljust(2)[:2].encode()
    head += struct.pack(">HHBBBBH", t.year, t.timetuple().tm_yday, t.hour, t.minute, t.second, 0,
                        t.microsecond // 100)
    head += struct.pack(">HhhBBBBiHH", n_samples, fact, mult, 0, 0, 0, 1, 0, DATA_OFFSET, 48)
    head += struct.pack(">HHBBBB", 1000, 0, encoding_code, 1, RECORD_LENGTH.bit_length() - 1, 0)
    return head.ljust(DATA_OFFSET, b"\x00")


def encode_trace(tr: Trace, encoding="STEIM2") -> bytes:
    """All records of one trace (record sequence numbers restart at 1, like libmseed per trace)."""
    code, version = _ENCODINGS[encoding]
    n_frames = (RECORD_LENGTH - DATA_OFFSET) // 64
    out, i, seq = [], 0, 1
    while i < tr.npts:
        d0 = 0 if i == 0 else int(tr.data[i]) - int(tr.data[i - 1])
        # a record never holds more than 63 * 15 * 7 samples: bound the slice handed to C
        chunk = tr.data[i:i + n_frames * 105]
        body, n = steim_encode(chunk, d0, version, n_frames)
        out.append(_pack_header(seq, tr, tr.time_of(i), n, code) + body)
        i += n
        seq += 1
    return b"".join(out)


def write_mseed(path, traces, encoding="STEIM2") -> None:
    """``Stream.write(path, format="MSEED", encoding=...)``: traces in order, 4096-byte records."""
    blob = b"".join(encode_trace(tr, encoding) for tr in traces if tr.npts)
    with open(path, "wb") as f:
        f.write(blob)


def _parse_records(blob: bytes):
    off, records = 0, []
    while off + 48 <= len(blob):
        head = blob[off:off + 48]
        if not head[:6].isdigit() or head[6:7] not in (b"D", b"R", b"Q", b"M"):
            raise ValueError(f"not a miniSEED record at byte {off}")
        year, day, hour, minute, sec, _, fract = struct.unpack(">HHBBBBH", head[20:30])
        n_samples, fact, mult = struct.unpack(">Hhh", head[30:36])
        n_blk, _tc, data_off, blk_off = struct.unpack(">BiHH", head[39:48])
        encoding, reclen, byteorder = None, RECORD_LENGTH, 1
        nxt, guard = blk_off, 0
        while nxt and guard < 16:
            btype, bnext = struct.unpack(">HH", blob[off + nxt:off + nxt + 4])
            if btype == 1000:
                encoding, byteorder, exp = struct.unpack(">BBB", blob[off + nxt + 4:off + nxt + 7])
                reclen = 1 << exp
            nxt, guard = bnext, guard + 1
        if encoding not in (10, 11) or byteorder != 1:
            raise ValueError(f"unsupported miniSEED encoding {encoding} / byte order {byteorder}")
        start = _dt.datetime(year, 1, 1) + _dt.timedelta(days=day - 1, hours=hour, minutes=minute,
                                                         seconds=sec, microseconds=fract * 100)
        rec = {
            "id": (head[18:20].decode().strip(), head[8:13].decode().strip(),
                   head[13:15].decode().strip(), head[15:18].decode().strip()),
            "start_us": to_us(start), "n": n_samples, "sr": _rate_from_factors(fact, mult),
            "version": 2 if encoding == 11 else 1,
            "frames": blob[off + data_off:off + reclen],
        }
        records.append(rec)
        off += reclen
    return records


def read_mseed(path, starttime=None, endtime=None):
    """
    ``obspy.read(path, format="MSEED", starttime=, endtime=)`` followed by ``merge(method=-1)`` for
    the subset of miniSEED the reference writes: records of one id that follow each other within
    half a sample are joined, anything else starts a new trace.  Traces come back in file order.
    """
    with open(path, "rb") as f:
        blob = f.read()
    traces, order = {}, []
    for rec in _parse_records(blob):
        if rec["n"] == 0:
            continue
        data, _ = steim_decode(rec["frames"], rec["n"], rec["version"])
        net, sta, loc, cha = rec["id"]
        tr = Trace(data, sta, rec["start_us"], rec["sr"], net, loc, cha)
        segs = traces.setdefault(rec["id"], [])
        if rec["id"] not in order:
            order.append(rec["id"])
        if segs:
            last = segs[-1]
            expect = last.time_of(last.npts)
            if abs(tr.start_us - expect) <= 0.5 * last.delta_us and last.sampling_rate == tr.sampling_rate:
                last.extend(tr.data)
                continue
        segs.append(tr)
    out = []
    s_us = to_us(starttime) if starttime is not None else None
    e_us = to_us(endtime) if endtime is not None else None
    for key in order:
        for tr in traces[key]:
            if s_us is not None or e_us is not None:
                tr = tr.slice(s_us, e_us)
            if tr.npts:
                out.append(tr)
    return out


# ------------------------------------------------------------------------------------------------
# ScanmSEED
# ------------------------------------------------------------------------------------------------
def _run_path(run) -> pathlib.Path:
    return pathlib.Path(getattr(run, "path", run))


class ScanmSEED:
    """
    Coalescence stream of ``detect()``: channels ``COA, COA_N, X, Y, Z`` as int32
    (scanmseed.py:23-72).  ``run`` is the run directory (or any object with a ``.path``).
    """

    def __init__(self, run, continuous_write, sampling_rate):
        self.run = run
        self.continuous_write = continuous_write
        self.sampling_rate = sampling_rate
        self.written = False
        self.stream: list[Trace] = []

    # -- scanmseed.py:226-240 -----------------------------------------------------------------
    @staticmethod
    def _data2int(data, factor):
        return np.round(np.asarray(data) * factor).astype(np.int32)

    def append(self, starttime, max_coa, max_coa_n, coord, ucf):
        """scanmseed.py:74-150 (the in-place clip of the caller's arrays included)."""
        max_coa[max_coa > 21474.0] = 21474.0
        max_coa_n[max_coa_n > 21474.0] = 21474.0
        self.append_quantised(
            starttime,
            self._data2int(max_coa, 1e5),
            self._data2int(max_coa_n, 1e5),
            self._data2int(coord[:, 0], 1e6),
            self._data2int(coord[:, 1], 1e6),
            self._data2int(coord[:, 2], 1e3 * ucf),
        )

    def append_from_engine(self, starttime, engine, n_samples, lut, ucf=None):
        """
        Same as :meth:`append`, with COA / COA_N / Z quantised on the device and only int32 channels
        (plus the float64 grid coordinates the projection needs) copied back.  ``lut`` supplies
        ``node_count``, ``ll_corner``, ``node_spacing``, ``unit_conversion_factor`` and the
        projection (``grid2coord``, lut.py:236).
        """
        ucf = lut.unit_conversion_factor if ucf is None else ucf
        q = engine.detect_fetch_quantised(n_samples, lut.node_count, lut.ll_corner, lut.node_spacing,
                                          1e3 * ucf)
        xyz = q["xyz"]
        coord = lut.grid2coord(xyz) if getattr(lut, "grid2coord", None) is not None else xyz
        self.append_quantised(starttime, q["coa"], q["coa_n"], self._data2int(coord[:, 0], 1e6),
                              self._data2int(coord[:, 1], 1e6), q["z"])
        return q

    def append_quantised(self, starttime, coa, coa_n, x, y, z):
        start_us = to_us(starttime)
        new = [Trace(d, sta, start_us, self.sampling_rate, "NW")
               for sta, d in zip(_CHANNELS, (coa, coa_n, x, y, z))]
        self._merge(new)

        # write to file if passed day line (scanmseed.py:131-147)
        self.written = False
        first = self.stream[0]
        midnight = _day_start_us(first.start_us) + 86_400_000_000
        delta = int(round(first.delta_us))
        if first.end_us == midnight - delta:
            self.write()
            self.stream = []
        elif _julday(first.start_us) != _julday(first.end_us):
            write_start = first.start_us
            write_end = _day_start_us(first.end_us) - delta
            self.write(write_start, write_end)
            self.stream = [t for t in (tr.slice(write_end + delta, None) for tr in self.stream) if t.npts]
            self.written = False

        if self.continuous_write and not self.written:
            self.write()

    def _merge(self, new):
        """``Stream.merge(method=-1)``: join traces of one id that are exactly contiguous."""
        for tr in new:
            for old in reversed(self.stream):
                if old.id == tr.id:
                    if abs(old.time_of(old.npts) - tr.start_us) <= 0.5 * old.delta_us:
                        old.extend(tr.data)
                        tr = None
                    break
            if tr is not None:
                self.stream.append(tr)
        # keep the channel order of the reference stream (COA first)
        self.stream.sort(key=lambda t: (_CHANNELS.index(t.station), t.start_us))

    def empty(self, starttime, timestep, i, msg, ucf):
        """scanmseed.py:152-185: a timestep without data is stored as zeros."""
        start_us = to_us(starttime) + int(round(timestep * i * 1e6))
        n = int(round(timestep * self.sampling_rate))  # util.time2sample (util.py:152-172)
        zeros = np.zeros(n, dtype=np.int32)
        self.append_quantised(start_us, zeros, zeros, zeros, zeros, zeros)

    def write(self, write_start=None, write_end=None):
        """scanmseed.py:187-224: ``<run>/detect/scanmseed/<year>_<julday>.scanmseed``."""
        fpath = _run_path(self.run) / "detect" / "scanmseed"
        fpath.mkdir(exist_ok=True, parents=True)
        if write_start is not None and write_end is not None:
            st = [t for t in (tr.slice(to_us(write_start), to_us(write_end)) for tr in self.stream) if t.npts]
        else:
            st = self.stream
        if not st:
            return None
        year, julday = _julday(st[0].start_us)
        file = fpath / f"{year}_{julday:03d}.scanmseed"
        try:
            write_mseed(file, st, encoding="STEIM2")
        except _lib.QmbError as e:
            if e.status != 3:  # QMB_ERR_UNSUPPORTED: a difference beyond 30 bits
                raise
            write_mseed(file, st, encoding="STEIM1")  # scanmseed.py:219-224
        self.written = True
        return file


def read_scanmseed(run, starttime, endtime, pad, ucf):
    """
    scanmseed.py:243-327.  Returns ``(data, stats)``: a ``pandas.DataFrame`` with columns
    ``DT, COA, COA_N, X, Y, Z`` (scale factors removed) and a dict with the COA trace's
    ``network, station, starttime, endtime, sampling_rate, delta, npts``.
    """
    import pandas as pd

    fpath = _run_path(run) / "detect" / "scanmseed"
    readstart, readend = to_us(starttime) - int(round(pad * 1e6)), to_us(endtime) + int(round(pad * 1e6))
    startday = _day_start_us(readstart)

    dy, traces = 0, []
    while startday + dy * 86_400_000_000 <= readend:
        year, julday = _julday(readstart + dy * 86_400_000_000)
        file = fpath / f"{year}_{julday:03d}.scanmseed"
        if file.exists():
            traces += read_mseed(file, starttime=readstart, endtime=readend)
        dy += 1
    if not traces:
        raise NoScanMseedDataException

    merged = ScanmSEED(run, False, traces[0].sampling_rate)
    merged._merge(traces)
    by_station = {}
    for tr in merged.stream:
        by_station.setdefault(tr.station, tr)  # first (earliest) trace of each channel
    coa = by_station["COA"]
    data = pd.DataFrame()
    data["DT"] = pd.to_datetime(merged.stream[0].times_us(), unit="us")
    data["COA"] = by_station["COA"].data / 1e5
    data["COA_N"] = by_station["COA_N"].data / 1e5
    data["X"] = by_station["X"].data / 1e6
    data["Y"] = by_station["Y"].data / 1e6
    data["Z"] = by_station["Z"].data / (1e3 * ucf)
    stats = {"network": coa.network, "station": coa.station, "starttime": coa.starttime,
             "endtime": coa.endtime, "sampling_rate": coa.sampling_rate,
             "delta": 1.0 / coa.sampling_rate, "npts": coa.npts}
    return data, stats
