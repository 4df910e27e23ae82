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
``channels`` (dict ``station -> {component letter -> 1-D float array}``) works;
:class:`WaveformData` is the minimal one.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .synthetic import time2sample, trim2sample


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


class WaveformData:
    """Minimal stand-in for ``quakemigrate.io.data.WaveformData`` (pre-processed data only)."""

    def __init__(self, channels: dict, starttime: float, sampling_rate: int):
        self.channels = channels
        self.stations = list(channels.keys())
        self.starttime = starttime
        self.sampling_rate = sampling_rate
        n = max((len(v) for comp in channels.values() for v in comp.values()), default=0)
        self.endtime = starttime + (n - 1) / sampling_rate if n else starttime


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
    def _plan(self, data):
        """
        Row order and availability exactly like stalta.py:389-470: phase-outer, station-inner,
        key ``f"{station}_{phase}"``; a row is available when at least one (``all_channels``:
        every) mapped component is present with the full length.
        """
        n_total = None
        availability, rows, chan_arrays = {}, [], []
        chan_index = {}
        per_phase = self.bandpass_filters is not None  # filtered traces are specific to a phase
        for phase in self.phases:
            comps = self.channel_maps[phase]
            stw, ltw = self.sta_lta_windows[phase]
            nsta = time2sample(stw, self.sampling_rate) + 1  # stalta.py:395-396
            nlta = time2sample(ltw, self.sampling_rate) + 1
            for station in data.stations:
                have = [c for c in comps if c in data.channels.get(station, {})]
                ok = len(have) == len(comps) if self.all_channels else len(have) > 0
                if ok:
                    lengths = {len(data.channels[station][c]) for c in have}
                    n_here = lengths.pop() if len(lengths) == 1 else None
                    ok = n_here is not None and (n_total is None or n_here == n_total)
                    if ok and n_total is None:
                        n_total = n_here
                availability[f"{station}_{phase}"] = 1 if ok else 0
                if not ok:
                    continue
                idxs = []
                for c in have:
                    key = (station, c, phase) if per_phase else (station, c)
                    if key not in chan_index:
                        chan_index[key] = len(chan_arrays)
                        chan_arrays.append(np.asarray(data.channels[station][c], dtype=np.double))
                    idxs.append(chan_index[key])
                rows.append((station, phase, idxs, nsta, nlta))
        self._last_chan_index = chan_index
        return availability, rows, chan_arrays

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
        waveforms = np.stack(chan_arrays, axis=0)
        filtered_waveforms = data.channels
        if self.bandpass_filters is not None:
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
