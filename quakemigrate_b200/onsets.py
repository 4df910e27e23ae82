# -*- coding: utf-8 -*-
"""
Onset plugin API of the coalescence hot path, mirrored from the reference so that code written
against ``quakemigrate.signal.onsets`` keeps working:

* :class:`Onset` -- the abstract base (``quakemigrate/signal/onsets/base.py:24-130``): same
  constructor contract (``sampling_rate`` required), same ``pad(timespan)`` arithmetic, same
  abstract ``calculate_onsets`` / ``pre_pad`` / ``post_pad``.
* :class:`OnsetData` -- the light result container (``base.py:133-191``), plus an optional
  ``device_onsets`` handle so that the scan can skip a second host->device copy.
* :class:`STALTAOnset` -- the default onset (``quakemigrate/signal/onsets/stalta.py:214-635``)
  with the per-channel STA/LTA, RMS combine and clip running on the GPU
  (``qmb_onsets_from_waveforms``).  With ``bandpass_filters`` the pre-processing of
  ``stalta.py:137-211`` (detrend, taper, zero-phase Butterworth band-pass per phase) also runs on
  the device and the filtered traces feed the STA/LTA kernels without a host round trip;
  without it, waveforms are taken as already pre-processed.

Waveform containers are duck-typed: anything with ``stations``, ``starttime``, ``endtime`` and
``channels`` (dict ``station -> {component letter -> trace}``) works; :class:`WaveformData` is the
minimal one.  A trace is a 1-D float array covering the whole timespan, or -- for gappy / short
data -- a list of :class:`Segment` (sample offset from ``starttime``, data).  Availability follows
``WaveformData.check_availability`` (io/data.py:509-646): flat-lined, overlapping, gappy (unless
``allow_gaps``) or short (if ``full_timespan``) channels are dropped; with ``allow_gaps`` or
``full_timespan=False`` the surviving segments are pre-processed one by one, tapered again, merged
and padded with tiny values exactly like stalta.py:436-462.
:mod:`quakemigrate_b200.reference_adapter` converts the reference's obspy-backed container.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .util import time2sample, trim2sample


class Onset(ABC):
    """QuakeMigrate onset base class (base.py:24-130)."""

    def __init__(self, **kwargs):
        self.sampling_rate = kwargs.get("sampling_rate")
        if self.sampling_rate is None:
            raise ValueError("Must specify 'sampling_rate' for any Onset.")
        self._pre_pad = 0
        self._post_pad = 0

    def __str__(self):
        return "Base Onset object - add a __str__ method to your Onset class"

    def pad(self, timespan):
        """Pre-/post-pad for a timespan including the 6 % taper allowance (base.py:64-93)."""
        timespan += self.pre_pad + self.post_pad
        pre_pad = trim2sample(self.pre_pad + np.ceil(timespan * 0.06), self.sampling_rate)
        post_pad = trim2sample(self.post_pad + np.ceil(timespan * 0.06), self.sampling_rate)
        return pre_pad, post_pad

    def gaussian_halfwidth(self, phase):
        raise AttributeError(
            "In order to use the GaussianPicker module with a custom Onset, you need "
            "to provide a 'gaussian_halfwidth' method."
        )

    @abstractmethod
    def calculate_onsets(self):
        """Return ``(onsets ndarray (n_onsets, n_samples) float64, OnsetData)``."""

    @property
    @abstractmethod
    def pre_pad(self):
        return self._pre_pad

    @pre_pad.setter
    @abstractmethod
    def pre_pad(self, value):
        self._pre_pad = value

    @property
    @abstractmethod
    def post_pad(self):
        return self._post_pad

    @post_pad.setter
    @abstractmethod
    def post_pad(self, value):
        self._post_pad = value


class OnsetData:
    """Result container of an onset calculation (base.py:133-191)."""

    def __init__(self, onsets, phases, channel_maps, filtered_waveforms, availability, starttime,
                 endtime, sampling_rate, device_onsets=None):
        self.onsets = onsets
        self.phases = phases
        self.channel_maps = channel_maps
        self.filtered_waveforms = filtered_waveforms
        self.availability = availability
        self.starttime = starttime
        self.endtime = endtime
        self.sampling_rate = sampling_rate
        # optional quakemigrate_b200.core.DeviceOnsets with the same rows already on the GPU
        self.device_onsets = device_onsets


class FilteredWaveforms:
    """
    The pre-processed traces of an onset calculation (``OnsetData.filtered_waveforms``; the
    reference keeps an obspy Stream, stalta.py:470), left on the device and copied out on demand:
    the picker and the plots want them, detect() never does.
    """

    def __init__(self, device_waveforms, index):
        self.device, self.index = device_waveforms, index  # {(station, component, phase): row}

    def trace(self, station, component, phase):
        return self.device.fetch(self.index[(station, component, phase)], 1)[0]

    def to_host(self):
        block = self.device.fetch()
        return {key: block[i] for key, i in self.index.items()}


class Segment:
    """One contiguous piece of a trace: ``offset`` samples after the container's ``starttime``."""

    __slots__ = ("offset", "data")

    def __init__(self, offset: int, data):
        self.offset, self.data = int(offset), np.asarray(data, dtype=np.double)

    def __len__(self):
        return len(self.data)


def _as_segments(trace):
    """A trace as a list of segments sorted by offset, perfectly contiguous neighbours merged."""
    if isinstance(trace, Segment):
        segs = [trace]
    elif isinstance(trace, (list, tuple)) and (not trace or isinstance(trace[0], Segment)):
        segs = sorted(trace, key=lambda g: g.offset)
    else:
        return [Segment(0, trace)]
    merged = []
    for g in segs:
        if merged and merged[-1].offset + len(merged[-1]) == g.offset:
            merged[-1] = Segment(merged[-1].offset, np.concatenate([merged[-1].data, g.data]))
        else:
            merged.append(g)
    return merged


class WaveformData:
    """Minimal stand-in for ``quakemigrate.io.data.WaveformData`` (data at the onset's sampling rate)."""

    def __init__(self, channels: dict, starttime: float, sampling_rate: int, endtime: float | None = None):
        self.channels = channels
        self.stations = list(channels.keys())
        self.starttime = starttime
        self.sampling_rate = sampling_rate
        if endtime is None:
            n = 0
            for comps in channels.values():
                for tr in comps.values():
                    n = max([n] + [g.offset + len(g) for g in _as_segments(tr)])
            endtime = starttime + (n - 1) / sampling_rate if n else starttime
        self.endtime = endtime


class DataAvailabilityException(Exception):
    """No onset function could be computed (quakemigrate/util.py DataAvailabilityException)."""

    def __init__(self):
        super().__init__("DataAvailabilityException: no data available that satisfies the requirements")


class STALTAOnset(Onset):
    """
    STA/LTA onset (classic or centred) computed on the GPU.

    Keyword arguments follow the reference (stalta.py:214-350): ``position`` ("classic" |
    "centred"), ``phases``, ``sta_lta_windows`` ({phase: [sta_s, lta_s]}), ``channel_maps``
    ({phase: component letters, e.g. {"P": "Z", "S": "NE"}}), ``signal_transform``
    ("energy" | "abs" | "env" | "env_squared"; the envelopes are ``|hilbert(x)|`` via batched FFTs
    on the device), ``min_onset_value`` (default 0.4), ``all_channels``, and
    ``bandpass_filters`` ({phase: [lowcut_hz, highcut_hz, corners]}, stalta.py:279-283).  The
    reference always filters; here ``bandpass_filters=None`` (default) means "the container holds
    pre-processed traces", anything else runs ``pre_process`` on the device first.
    """

    def __init__(self, engine=None, **kwargs):
        super().__init__(**kwargs)
        self.engine = engine
        self.position = kwargs.get("position", "classic")
        if self.position not in ("classic", "centred"):
            raise ValueError("position must be 'classic' or 'centred'")
        self.phases = kwargs.get("phases", ["P", "S"])
        self.sta_lta_windows = kwargs.get("sta_lta_windows", {"P": [0.2, 1.0], "S": [0.2, 1.0]})
        self.channel_maps = kwargs.get("channel_maps", {"P": "Z", "S": "NE"})
        self.signal_transform = kwargs.get("signal_transform", "energy")
        if self.signal_transform not in ("energy", "abs", "env", "env_squared"):  # stalta.py:516-523
            raise ValueError(f"unknown signal_transform {self.signal_transform!r}")
        self.min_onset_value = kwargs.get("min_onset_value", 0.4)
        self.all_channels = kwargs.get("all_channels", False)
        self.channel_counts = kwargs.get("channel_counts", {"P": 1, "S": 2})  # stalta.py:322
        self.allow_gaps = kwargs.get("allow_gaps", False)                      # stalta.py:326
        self.full_timespan = kwargs.get("full_timespan", True)                 # stalta.py:327
        self.bandpass_filters = kwargs.get("bandpass_filters", None)
        if self.bandpass_filters is not None:
            from .preprocess import NyquistException
            for phase in self.phases:
                _, highcut, _ = self.bandpass_filters[phase]
                if highcut >= 0.5 * self.sampling_rate:  # stalta.py:198-199
                    raise NyquistException(highcut, 0.5 * self.sampling_rate, "")

    def __str__(self):
        out = (f"\tOnset parameters - using the {self.position} STA/LTA onset (GPU)"
               f"\n\t\tOnset function sampling rate = {self.sampling_rate} Hz"
               f"\n\t\tPhase(s) = {self.phases}\n")
        for phase, windows in self.sta_lta_windows.items():
            out += f"\n\t\t{phase} onset [STA, LTA] = {windows} (s, s)"
        return out + "\n"

    # -- pads (stalta.py:600-635) -------------------------------------------------------------
    @property
    def pre_pad(self):
        windows = self.sta_lta_windows
        return max(w[1] for w in windows.values()) + 3 * max(w[0] for w in windows.values())

    @pre_pad.setter
    def pre_pad(self, value):
        self._pre_pad = value

    @property
    def post_pad(self):
        return self._post_pad

    @post_pad.setter
    def post_pad(self, ttmax):
        lta_max = max(w[1] for w in self.sta_lta_windows.values())
        self._post_pad = np.ceil(ttmax + 2 * lta_max)

    def gaussian_halfwidth(self, phase):
        return self.sta_lta_windows[phase][0] * self.sampling_rate / 2

    # -- the onset calculation ------------------------------------------------------------------
    def _expected_samples(self, data) -> int:
        """Samples a trace covering the container's timespan holds (io/data.py:612: timespan * sr + 1)."""
        return int(round(float(data.endtime - data.starttime) * self.sampling_rate)) + 1

    def _channel_ok(self, segs, n_expected) -> bool:
        """One channel against the criteria of ``check_availability`` (io/data.py:585-633)."""
        if not segs or any(len(g) == 0 for g in segs):
            return False
        if any(g.data.max() == g.data.min() for g in segs):          # flat-lined (:594-595)
            return False
        ends = [g.offset + len(g) for g in segs]
        if any(b.offset < e for e, b in zip(ends[:-1], segs[1:])):   # overlaps (:597-599)
            return False
        if not self.allow_gaps and len(segs) > 1:                    # gaps (:601-604)
            return False
        if self.full_timespan and (len(segs) > 1 or len(segs[0]) < n_expected):   # (:611-621)
            return False
        return True

    def _plan(self, data):
        """
        Row order and availability exactly like stalta.py:389-470: phase-outer, station-inner,
        key ``f"{station}_{phase}"``.  A channel counts when it passes :meth:`_channel_ok`; a row is
        available when at least one mapped component does (``all_channels``: every one of them, and
        as many as ``channel_counts[phase]``, io/data.py:636-650).  The expected trace length comes
        from the container's timespan, not from whichever trace happens to be seen first.
        """
        n_expected = self._expected_samples(data)
        strict = self.full_timespan and not self.allow_gaps
        availability, rows, chan_arrays = {}, [], []
        chan_index = {}
        per_phase = self.bandpass_filters is not None  # filtered traces are specific to a phase
        for phase in self.phases:
            comps = self.channel_maps[phase]
            stw, ltw = self.sta_lta_windows[phase]
            nsta = time2sample(stw, self.sampling_rate) + 1  # stalta.py:395-396
            nlta = time2sample(ltw, self.sampling_rate) + 1
            for station in data.stations:
                present = {c: _as_segments(data.channels[station][c])
                           for c in comps if c in data.channels.get(station, {})}
                good = {c: segs for c, segs in present.items() if self._channel_ok(segs, n_expected)}
                if strict:  # a full-length trace must also start on the first sample
                    good = {c: segs for c, segs in good.items()
                            if segs[0].offset == 0 and len(segs[0]) == n_expected}
                if self.all_channels:
                    ok = len(present) > 0 and len(good) == len(present) == self.channel_counts.get(phase, len(comps))
                else:
                    ok = len(good) > 0
                availability[f"{station}_{phase}"] = 1 if ok else 0
                if not ok:
                    continue
                idxs = []
                for c, segs in good.items():
                    key = (station, c, phase) if per_phase else (station, c)
                    if key not in chan_index:
                        chan_index[key] = len(chan_arrays)
                        chan_arrays.append(segs[0].data if strict else segs)
                    idxs.append(chan_index[key])
                rows.append((station, phase, idxs, nsta, nlta))
        self._last_chan_index = chan_index
        return availability, rows, chan_arrays

    def _assemble_gappy(self, chan_segments, chan_index, n_expected):
        """
        stalta.py:436-462 for ``allow_gaps`` / ``full_timespan=False``: every segment is
        pre-processed on its own (the obspy calls of ``pre_process`` work trace by trace), tapered
        once more (5 % cosine), then the segments of a channel are merged, gaps and the missing
        start / end filled with ``sqrt(tiny)`` so that the energy trace stays positive.
        Returns full-length float64 channel arrays ``(n_channels, n_expected)``.
        """
        from .preprocess import butter_sos, cosine_taper_window

        tiny = np.sqrt(np.finfo(float).tiny)
        phase_of = {i: (key[2] if len(key) == 3 else None) for key, i in chan_index.items()}
        out = np.full((len(chan_segments), n_expected), tiny)
        sos_cache = {}
        for i, segs in enumerate(chan_segments):
            phase = phase_of[i]
            for g in segs:
                x = g.data
                if self.bandpass_filters is not None:
                    if phase not in sos_cache:
                        lowcut, highcut, corners = self.bandpass_filters[phase]
                        sos_cache[phase] = butter_sos("bandpass", (lowcut, highcut), self.sampling_rate, corners)
                    x = self.engine.preprocess(x[None, :], sos_cache[phase])[0]
                x = x * cosine_taper_window(len(x), 0.05)
                a, b = max(g.offset, 0), min(g.offset + len(x), n_expected)   # trim to the timespan (:454-460)
                if b > a:
                    out[i, a:b] = x[a - g.offset:b - g.offset]
        return out

    def _filter_on_device(self, waveforms, chan_index):
        """``pre_process`` (stalta.py:203-209) per phase; returns the filtered DeviceWaveforms."""
        from .preprocess import butter_sos

        raw = self.engine.upload_waveforms(waveforms)
        try:
            filtered = self.engine.alloc_waveforms(*raw.shape)
            for phase in self.phases:
                sel = [i for key, i in chan_index.items() if key[2] == phase]
                if not sel:
                    continue
                lowcut, highcut, corners = self.bandpass_filters[phase]
                sos = butter_sos("bandpass", (lowcut, highcut), self.sampling_rate, corners)
                # channels of one phase are contiguous in first-seen order: filter them in runs
                start = 0
                while start < len(sel):
                    stop = start + 1
                    while stop < len(sel) and sel[stop] == sel[stop - 1] + 1:
                        stop += 1
                    self.engine.filter_waveforms(raw, sos, channels=sel[start:stop], out=filtered,
                                                 out_chan0=sel[start])
                    start = stop
            return filtered
        finally:
            raw.free()

    def calculate_onsets(self, data, timespan=None):
        """
        Returns ``(onsets, onset_data)`` like stalta.py:353-489.  ``onsets`` is float64
        ``(n_available_rows, n_samples)``; ``onset_data.device_onsets`` holds the same rows
        (clip+log applied) on the GPU for the scan.
        """
        if self.engine is None:
            from .core import Engine
            self.engine = Engine(0)
        availability, rows, chan_arrays = self._plan(data)
        if sum(availability.values()) == 0:
            raise DataAvailabilityException
        gappy = self.allow_gaps or not self.full_timespan
        if gappy:
            waveforms = self._assemble_gappy(chan_arrays, self._last_chan_index, self._expected_samples(data))
        else:
            waveforms = np.stack(chan_arrays, axis=0)
        filtered_waveforms = data.channels
        if gappy and self.bandpass_filters is not None:
            # already filtered segment by segment: only the onset stage is left
            raw = self.engine.upload_waveforms(waveforms)
            dev, onsets = self.engine.onsets_from_device_waveforms(
                raw, [r[2] for r in rows], [r[3] for r in rows], [r[4] for r in rows],
                position=self.position, transform=self.signal_transform,
                min_onset_value=self.min_onset_value, return_onsets=True)
            filtered_waveforms = FilteredWaveforms(raw, dict(self._last_chan_index))
        elif self.bandpass_filters is not None:
            filtered = self._filter_on_device(waveforms, self._last_chan_index)
            dev, onsets = self.engine.onsets_from_device_waveforms(
                filtered, [r[2] for r in rows], [r[3] for r in rows], [r[4] for r in rows],
                position=self.position, transform=self.signal_transform,
                min_onset_value=self.min_onset_value, return_onsets=True)
            filtered_waveforms = FilteredWaveforms(filtered, dict(self._last_chan_index))
        else:
            dev, onsets = self.engine.onsets_from_waveforms(
                waveforms, [r[2] for r in rows], [r[3] for r in rows], [r[4] for r in rows],
                position=self.position, transform=self.signal_transform,
                min_onset_value=self.min_onset_value, return_onsets=True)
        if timespan:
            # _trim_taper_pad (stalta.py:550-584): the tapered pre-/post-pad plus one LTA / STA
            # window is set to 1 on every channel before the RMS combine; every channel of a row
            # shares its windows, so RMS(1, .., 1) = 1 >= min_onset_value and nulling the combined
            # row is identical.  Only the picker asks for this (gaussian.py:142): the device copy
            # is dropped so that nobody migrates nulled onsets by accident.
            pre_pad, _ = self.pad(timespan)
            taper_pad = time2sample(pre_pad - self.pre_pad, self.sampling_rate)
            for i, (_, _, _, nsta, nlta) in enumerate(rows):
                onsets[i, : (taper_pad + nlta - 1)] = 1.0
                onsets[i, -(nsta + taper_pad):] = 1.0
            dev.free()
            dev = None
        onsets_dict = {}
        for i, (station, phase, _, _, _) in enumerate(rows):
            onsets_dict.setdefault(station, {})[phase] = onsets[i]
        onset_data = OnsetData(onsets_dict, self.phases, self.channel_maps, filtered_waveforms, availability,
                               data.starttime, data.endtime, self.sampling_rate, device_onsets=dev)
        return onsets, onset_data
