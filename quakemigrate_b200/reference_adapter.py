# -*- coding: utf-8 -*-
"""
Plugging quakemigrate_b200 into an UNMODIFIED ``quakemigrate`` installation.

The reference's ``QuakeScan`` refuses any onset that is not an instance of its own ``Onset`` ABC
(``quakemigrate/signal/scan.py:209-212``) and hands ``calculate_onsets`` an obspy-backed
``WaveformData`` (``quakemigrate/io/data.py:390``: ``data.waveforms`` is an obspy ``Stream``,
consumed at ``quakemigrate/signal/onsets/stalta.py:353-489``).  This module closes both gaps:

* :func:`register_with_reference` registers our :class:`~quakemigrate_b200.onsets.Onset` as a
  virtual subclass of ``quakemigrate.signal.onsets.base.Onset`` (``ABCMeta.register``), so every
  onset defined here passes the ``isinstance`` check;
* :func:`container_from_stream` turns an obspy ``Stream`` (anything iterable whose items carry
  ``.data`` and ``.stats.{station, channel, starttime, sampling_rate, npts}``) into the
  segment-based container of :mod:`quakemigrate_b200.onsets`: traces of one channel that obspy
  keeps as several pieces (gaps) become several :class:`~quakemigrate_b200.onsets.Segment`, traces
  at a multiple of the onset's sampling rate are decimated on the device like ``util.resample``
  (``quakemigrate/util.py:404-520``) does per trace;
* :class:`ReferenceSTALTAOnset` is :class:`~quakemigrate_b200.onsets.STALTAOnset` whose
  ``calculate_onsets`` accepts the reference's container (or ours);
* :func:`gpu_quakescan_class` builds, where ``quakemigrate`` is importable, the subclass of the
  reference's own ``QuakeScan`` whose ``_compute`` (scan.py:594-647) is the fused device path --
  everything else (archive reading, output files, picking, magnitudes, plots) stays the
  reference's code.

obspy / quakemigrate are imported lazily and only by the two functions that need them; nothing
here is required by the rest of the package.
"""

from __future__ import annotations

import numpy as np

from .onsets import Onset, STALTAOnset, Segment, WaveformData


def register_with_reference() -> bool:
    """``quakemigrate.signal.onsets.base.Onset.register(Onset)``; False when quakemigrate is absent."""
    try:
        from quakemigrate.signal.onsets.base import Onset as ReferenceOnset
    except ImportError:
        return False
    ReferenceOnset.register(Onset)
    return True


def _seconds(t) -> float:
    """UTCDateTime (``.timestamp``), datetime (``.timestamp()``) or plain number -> float seconds."""
    ts = getattr(t, "timestamp", None)
    if ts is None:
        return float(t)
    return float(ts() if callable(ts) else ts)


def container_from_stream(stream, stations, starttime, endtime, sampling_rate, engine=None,
                          component=lambda channel: channel[-1]) -> WaveformData:
    """
    obspy ``Stream`` -> :class:`~quakemigrate_b200.onsets.WaveformData` at ``sampling_rate``.

    Every trace becomes one segment of its (station, component); its offset is the number of
    samples between ``starttime`` and the trace's first sample (the reference shifts data onto the
    sample grid when reading, ``io/data.py:300-303``, so this is an integer).  A trace recorded at
    an integer multiple of ``sampling_rate`` is decimated first -- detrend, cosine taper,
    zero-phase low-pass at ``sampling_rate / 2.000001``, every ``factor``-th sample
    (``util.py:479-520``) -- on the device (``engine``); other rates are skipped like
    ``util.resample`` does without ``resample=True`` (``util.py:461-466``).
    """
    from . import preprocess as pp

    t0 = _seconds(starttime)
    channels: dict = {}
    for tr in stream:
        st = tr.stats
        if stations is not None and st.station not in stations:
            continue
        rate = float(st.sampling_rate)
        data = np.asarray(tr.data, dtype=np.double)
        if rate != sampling_rate:
            if rate % sampling_rate != 0 or len(data) < 2:
                continue
            data = pp.decimate(data[None, :], rate, sampling_rate, engine=engine)[0]
        offset = int(round((_seconds(st.starttime) - t0) * sampling_rate))
        channels.setdefault(st.station, {}).setdefault(component(st.channel), []).append(Segment(offset, data))
    out = WaveformData(channels, starttime, sampling_rate, endtime=endtime)
    out.stations = list(stations) if stations is not None else list(channels)
    return out


class ReferenceSTALTAOnset(STALTAOnset):
    """
    :class:`~quakemigrate_b200.onsets.STALTAOnset` that accepts ``quakemigrate.io.data.WaveformData``.

    ``channel_maps`` may be given in the reference's glob form (``{"P": "*Z", "S": "*[N,E,1,2]"}``,
    stalta.py:321): the component letters are taken from it.  ``data.waveforms`` is converted with
    :func:`container_from_stream`; a container of this package passes through untouched.
    """

    def __init__(self, engine=None, **kwargs):
        maps = kwargs.get("channel_maps")
        if maps is not None:
            kwargs = dict(kwargs, channel_maps={ph: self._letters(pattern) for ph, pattern in maps.items()})
        kwargs.setdefault("bandpass_filters", {"P": [2.0, 16.0, 2], "S": [2.0, 16.0, 2]})  # stalta.py:319
        super().__init__(engine=engine, **kwargs)

    @staticmethod
    def _letters(pattern: str) -> str:
        """'*Z' -> 'Z';  '*[N,E,1,2]' -> 'NE12';  'NE' -> 'NE'."""
        body = pattern.lstrip("*?")
        if body.startswith("[") and body.endswith("]"):
            body = body[1:-1]
        return "".join(ch for ch in body if ch.isalnum())

    def calculate_onsets(self, data, timespan=None):
        if hasattr(data, "waveforms") and not hasattr(data, "channels"):
            if self.engine is None:
                from .core import Engine
                self.engine = Engine(0)
            data = container_from_stream(data.waveforms, data.stations, data.starttime, data.endtime,
                                         self.sampling_rate, engine=self.engine)
        return super().calculate_onsets(data, timespan=timespan)


def gpu_quakescan_class():
    """
    ``class GpuQuakeScan(quakemigrate.signal.scan.QuakeScan)`` with ``_compute`` replaced by the
    fused device path.  Usage (an otherwise unmodified detect / locate script)::

        from quakemigrate_b200.reference_adapter import ReferenceSTALTAOnset, gpu_quakescan_class
        onset = ReferenceSTALTAOnset(position="classic", sampling_rate=50, ...)
        scan = gpu_quakescan_class()(archive, lut, onset=onset, run_path=..., run_name=...)
        scan.detect(starttime, endtime)
    """
    from quakemigrate.signal.scan import QuakeScan as ReferenceQuakeScan

    from .core import Engine
    from .scan import marginal_window_indices  # noqa: F401  (documented entry point for locate)

    register_with_reference()

    class GpuQuakeScan(ReferenceQuakeScan):
        def __init__(self, *args, engine=None, **kwargs):
            super().__init__(*args, **kwargs)
            self.engine = engine or getattr(self.onset, "engine", None) or Engine(0)
            if getattr(self.onset, "engine", "absent") is None:
                self.onset.engine = self.engine
            self._lut_cache = {}

        def _device_lut(self, sampling_rate, availability):
            key = (int(sampling_rate), tuple(availability.items()))
            if key not in self._lut_cache:
                if len(self._lut_cache) >= 4:
                    self._lut_cache.pop(next(iter(self._lut_cache))).free()
                self._lut_cache[key] = self.engine.upload_lut(
                    self.lut.serve_traveltimes(sampling_rate, availability))
            return self._lut_cache[key]

        def _compute(self, data, event=None):
            """scan.py:594-647 with migrate + find_max_coa fused on the device."""
            from quakemigrate import util

            onsets, onset_data = self.onset.calculate_onsets(data)
            try:
                dlut = self._device_lut(onset_data.sampling_rate, onset_data.availability)
            except KeyError as e:
                raise util.LUTPhasesException(
                    f"Attempting to migrate phases {onset_data.phases}; but traveltimes for {e} not found in "
                    f"the LUT. Please create a new lookup table with phases={onset_data.phases}"
                )
            fsmp = util.time2sample(self.pre_pad, onset_data.sampling_rate)
            lsmp = util.time2sample(self.post_pad, onset_data.sampling_rate)
            avail = int(np.sum([value for _, value in onset_data.availability.items()]))
            dev = getattr(onset_data, "device_onsets", None) or self.engine.upload_onsets(onsets)
            n_samples = dev.n_total - fsmp - lsmp
            self.engine.detect_device(dlut, dev, fsmp, lsmp, avail)
            max_coa, max_coa_n, max_idx = self.engine.detect_fetch(n_samples)
            coord = self.lut.index2coord(max_idx, unravel=True)
            if self.run.stage == "detect":
                return data.starttime + self.pre_pad, max_coa, max_coa_n, coord, onset_data
            # locate: the reference keeps the whole (nx, ny, nz, T) volume; T = 4 mw sr + 1 is small
            # here (event.py:398-419), so the window kernel materialises it once
            _, map4d = self.engine.locate_window(dlut, dev, fsmp, lsmp, avail, 0, n_samples, return_map4d=True)
            times = event.mw_times(self.scan_rate)
            return times, max_coa, max_coa_n, coord, map4d, onset_data

    return GpuQuakeScan
