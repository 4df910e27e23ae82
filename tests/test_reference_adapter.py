# -*- coding: utf-8 -*-
"""
The reference-side plug (VERDICT r1 missing #3 / #4): obspy-Stream -> segment container, gappy and
short traces handled like stalta.py:436-462 + io/data.py:509-646, and the ABC registration that
lets the reference's QuakeScan accept our onsets (scan.py:209-212).  obspy is absent here: the
Stream is a list of duck-typed traces, and the registration is exercised against a stand-in
``quakemigrate`` package with the reference's own Onset ABC restated.
"""
import sys
import types

import numpy as np
import pytest

from oracle import preprocess_oracle as PO
from oracle import qm_oracle as O
from quakemigrate_b200 import reference_adapter as ra
from quakemigrate_b200.onsets import STALTAOnset, Segment, WaveformData
from quakemigrate_b200.preprocess import butter_sos
from standin_engine import StandInEngine

SR = 50


class _Stats:
    def __init__(self, station, channel, starttime, sampling_rate, npts):
        self.station, self.channel, self.starttime = station, channel, starttime
        self.sampling_rate, self.npts = sampling_rate, npts


class _Trace:
    def __init__(self, station, channel, starttime, sampling_rate, data):
        self.data = np.asarray(data)
        self.stats = _Stats(station, channel, starttime, sampling_rate, len(self.data))


def _noise(n, seed):
    return np.random.default_rng(seed).standard_normal(n) * 100.0


def test_container_from_stream_offsets_gaps_and_decimation():
    n = 20 * SR + 1
    st = [
        _Trace("AAA", "HHZ", 1000.0, SR, _noise(n, 1)),                      # full
        _Trace("AAA", "HHN", 1000.0, SR, _noise(200, 2)),                    # two pieces: a gap
        _Trace("AAA", "HHN", 1000.0 + 300 / SR, SR, _noise(n - 300, 3)),
        _Trace("BBB", "EHZ", 1000.0, 2 * SR, _noise(2 * n - 1, 4)),          # 100 Hz -> decimated by 2
        _Trace("BBB", "EHE", 1000.0, 40, _noise(500, 5)),                    # 40 Hz: not a multiple -> skipped
        _Trace("CCC", "HHZ", 1000.0, SR, _noise(n, 6)),                      # station not requested
    ]
    eng = StandInEngine()
    data = ra.container_from_stream(st, ["AAA", "BBB"], 1000.0, 1000.0 + 20.0, SR, engine=eng)
    assert data.stations == ["AAA", "BBB"] and "CCC" not in data.channels
    assert [(g.offset, len(g)) for g in data.channels["AAA"]["N"]] == [(0, 200), (300, n - 300)]
    assert len(data.channels["BBB"]["Z"][0]) == n and "E" not in data.channels["BBB"]
    want = PO.decimate(st[3].data.astype(float), 2 * SR, SR)
    np.testing.assert_allclose(data.channels["BBB"]["Z"][0].data, want, rtol=1e-12, atol=1e-9)


def _onset(**kw):
    kw.setdefault("phases", ["P", "S"])
    return STALTAOnset(engine=StandInEngine(), sampling_rate=SR,
                       sta_lta_windows={"P": [0.2, 1.0], "S": [0.2, 1.0]}, **kw)


def _container(chan):
    d = WaveformData(chan, 0.0, SR, endtime=20.0)
    return d


def test_availability_rules_of_the_reference():
    n = 20 * SR + 1
    chan = {
        "FULL": {"Z": _noise(n, 1), "N": _noise(n, 2), "E": _noise(n, 3)},
        "GAPS": {"Z": [Segment(0, _noise(400, 4)), Segment(450, _noise(n - 450, 5))], "N": _noise(n, 6), "E": _noise(n, 7)},
        "SHRT": {"Z": _noise(n - 30, 8)},
        "FLAT": {"Z": np.full(n, 3.0)},
        "OVLP": {"Z": [Segment(0, _noise(500, 9)), Segment(480, _noise(n - 480, 10))]},
        "LATE": {"Z": [Segment(25, _noise(n - 25, 11))]},
    }
    data = _container(chan)
    # defaults (io/data.py:509-646 with allow_gaps=False, full_timespan=True)
    av, rows, _ = _onset()._plan(data)
    assert av == {"FULL_P": 1, "GAPS_P": 0, "SHRT_P": 0, "FLAT_P": 0, "OVLP_P": 0, "LATE_P": 0,
                  "FULL_S": 1, "GAPS_S": 1, "SHRT_S": 0, "FLAT_S": 0, "OVLP_S": 0, "LATE_S": 0}
    # allow_gaps + partial timespans: gappy / short / late channels come back, flat and overlapping never do
    av, rows, _ = _onset(allow_gaps=True, full_timespan=False)._plan(data)
    assert av["GAPS_P"] == av["SHRT_P"] == av["LATE_P"] == 1 and av["FLAT_P"] == av["OVLP_P"] == 0
    # all_channels: the S row needs both horizontals (channel_counts["S"] = 2)
    av, _, _ = _onset(all_channels=True)._plan(_container({"ONE": {"Z": _noise(n, 1), "N": _noise(n, 2)}}))
    assert av == {"ONE_P": 1, "ONE_S": 0}
    # the expected length comes from the timespan, not from the first trace seen
    av, _, _ = _onset()._plan(_container({"A": {"Z": _noise(n - 7, 1)}, "B": {"Z": _noise(n, 2)}}))
    assert av["A_P"] == 0 and av["B_P"] == 1


def test_gappy_channel_matches_the_reference_recipe():
    """stalta.py:401-462 restated with the oracle: per-segment pre_process, second 5 % taper, gaps and
    ends filled with sqrt(tiny), then the centred STA/LTA of the merged trace."""
    n = 20 * SR + 1
    segs = [Segment(30, _noise(400, 4)), Segment(520, _noise(n - 560, 5))]
    data = _container({"G": {"Z": segs}})
    filt = [2.0, 16.0, 2]
    on = _onset(allow_gaps=True, full_timespan=False, position="centred", phases=["P"],
                bandpass_filters={"P": filt})
    onsets, od = on.calculate_onsets(data)
    tiny = np.sqrt(np.finfo(float).tiny)
    full = np.full(n, tiny)
    for g in segs:
        x = PO.pre_process(g.data, SR, filt)
        x = x * PO.cosine_taper_window(len(x), 0.05)
        full[g.offset:g.offset + len(x)] = x
    want = np.clip(np.sqrt(O.centred_sta_lta(full ** 2, 11, 51) ** 2), 0.4, np.inf)
    assert od.availability == {"G_P": 1} and onsets.shape == (1, n)
    np.testing.assert_allclose(onsets[0], want, rtol=1e-9)


def test_reference_onset_accepts_the_obspy_container():
    n = 20 * SR + 1
    stream = [_Trace("AAA", "HHZ", 500.0, SR, _noise(n, 1)), _Trace("AAA", "HHN", 500.0, SR, _noise(n, 2)),
              _Trace("AAA", "HHE", 500.0, SR, _noise(n, 3))]
    ref_data = types.SimpleNamespace(waveforms=stream, stations=["AAA", "ZZZ"], starttime=500.0, endtime=520.0)
    on = ra.ReferenceSTALTAOnset(engine=StandInEngine(), sampling_rate=SR, position="classic",
                                 channel_maps={"P": "*Z", "S": "*[N,E,1,2]"})
    assert on.channel_maps == {"P": "Z", "S": "NE12"}
    onsets, od = on.calculate_onsets(ref_data)
    assert od.availability == {"AAA_P": 1, "ZZZ_P": 0, "AAA_S": 1, "ZZZ_S": 0} and onsets.shape == (2, n)
    sos = butter_sos("bandpass", (2.0, 16.0), SR, 2)
    z = PO.pre_process(stream[0].data.astype(float), SR, [2.0, 16.0, 2], sos=sos)
    np.testing.assert_allclose(onsets[0], np.clip(O.overlapping_sta_lta(z ** 2, 11, 51), 0.4, np.inf), rtol=1e-9)


def test_registration_passes_the_reference_isinstance_check(monkeypatch):
    """scan.py:209-212: ``if not isinstance(onset, Onset): raise TypeError``.  With a stand-in
    ``quakemigrate.signal.onsets.base`` holding an ABC like the reference's, registration makes our
    onsets instances of it."""
    from abc import ABC, abstractmethod

    class RefOnset(ABC):
        @abstractmethod
        def calculate_onsets(self):
            ...

    pkgs = {}
    for name in ("quakemigrate", "quakemigrate.signal", "quakemigrate.signal.onsets", "quakemigrate.signal.onsets.base"):
        pkgs[name] = types.ModuleType(name)
    pkgs["quakemigrate.signal.onsets.base"].Onset = RefOnset
    for name, mod in pkgs.items():
        monkeypatch.setitem(sys.modules, name, mod)
    on = _onset()
    assert not isinstance(on, RefOnset)
    assert ra.register_with_reference() is True
    assert isinstance(on, RefOnset)


def test_registration_without_quakemigrate_is_a_no_op(monkeypatch):
    monkeypatch.setitem(sys.modules, "quakemigrate", None)
    assert ra.register_with_reference() is False


@pytest.mark.gpu
def test_gappy_channel_on_the_device():
    """The same gappy recipe through libqmb200 (per-segment filter and the onset stage on the GPU)."""
    from quakemigrate_b200.core import Engine
    n = 20 * SR + 1
    segs = [Segment(30, _noise(400, 4)), Segment(520, _noise(n - 560, 5))]
    full_n = _noise(n, 6)
    data = _container({"G": {"Z": segs, "N": full_n, "E": [Segment(0, _noise(n - 100, 7))]}})
    filt = [2.0, 16.0, 2]
    with Engine(0) as eng:
        on = STALTAOnset(engine=eng, sampling_rate=SR, allow_gaps=True, full_timespan=False, position="centred",
                         phases=["P", "S"], sta_lta_windows={"P": [0.2, 1.0], "S": [0.2, 1.0]},
                         bandpass_filters={"P": filt, "S": filt})
        onsets, od = on.calculate_onsets(data)
        assert od.availability == {"G_P": 1, "G_S": 1} and onsets.shape == (2, n)
        assert od.device_onsets is not None
    tiny = np.sqrt(np.finfo(float).tiny)

    def assemble(pieces):
        full = np.full(n, tiny)
        for g in pieces:
            x = PO.pre_process(g.data, SR, filt)
            x = x * PO.cosine_taper_window(len(x), 0.05)
            full[g.offset:g.offset + len(x)] = x
        return full
    z = O.centred_sta_lta(assemble(segs) ** 2, 11, 51)
    hn = O.centred_sta_lta(assemble([Segment(0, full_n)]) ** 2, 11, 51)
    he = O.centred_sta_lta(assemble([Segment(0, _noise(n - 100, 7))]) ** 2, 11, 51)
    # at the edges of a gap the windows hold the sqrt(tiny) fill and a few tapered samples: the ratio is
    # ~1e7 and rounding-level differences of the filter show up at 1e-6 of it
    np.testing.assert_allclose(onsets[0], np.clip(z, 0.4, np.inf), rtol=2e-6)
    # The E trace ends 100 samples early.  Once BOTH windows of a sample lie in the sqrt(tiny) fill the
    # reference's sliding sums hold nothing but the rounding residue of the data that left them
    # (onsetlib.c:98-107 adds and subtracts), i.e. noise; the device sums come out as exact zeros and
    # the "lta <= 0 -> 1" rule applies.  Compare where at least the long window still sees data.
    ok = np.arange(n) < n - 100 + 40
    want_s = np.clip(np.sqrt((hn ** 2 + he ** 2) / 2), 0.4, np.inf)
    np.testing.assert_allclose(onsets[1][ok], want_s[ok], rtol=2e-6)
    assert np.isfinite(onsets[1]).all() and (onsets[1] >= 0.4).all()
