# -*- coding: utf-8 -*-
"""
``QuakeScan._compute`` on the GPU -- the single caller of the coalescence hot path in the
reference (``quakemigrate/signal/scan.py:594-647``), re-hosted with the same plugin contract:

    onsets, onset_data = self.onset.calculate_onsets(data)                    # scan.py:619
    traveltimes = self.lut.serve_traveltimes(sampling_rate, availability)     # scan.py:621-623
    fsmp, lsmp  = time2sample(pre_pad), time2sample(post_pad)                 # scan.py:632-633
    avail       = sum(availability.values())                                  # scan.py:634
    map4d       = migrate(onsets, traveltimes, fsmp, lsmp, avail, threads)    # scan.py:635
    max_coa, max_coa_n, max_idx = find_max_coa(map4d, threads)                # scan.py:638
    coord       = self.lut.index2coord(max_idx, unravel=True)                 # scan.py:639

Differences, all behind the same return values:

* migrate + find_max_coa are one fused device pass; ``map4d`` never exists in detect;
* the int32 travel-time stack is uploaded once per availability pattern and cached on the device
  (the reference rebuilds it every call, lut.py:502-538);
* locate does not return the full ``(nx, ny, nz, T)`` volume: it returns the marginal window the
  reference would cut out of it (``Event.trim2window``, io/event.py:421-439) and its time-sum
  (``coa_map``, scan.py:720) -- use :meth:`QuakeScan.compute_locate`.  The reference-shaped
  ``_compute(data, event)`` that materialises the whole volume is kept for small problems.

``onset`` may be any :class:`quakemigrate_b200.onsets.Onset` (or duck-typed equivalent returning
host arrays); if its ``OnsetData`` carries ``device_onsets`` the host->device copy is skipped.
"""

from __future__ import annotations

import logging

import numpy as np

from .core import DeviceOnsets, Engine
from .onsets import DataAvailabilityException, Onset
from .util import time2sample


class LUTPhasesException(Exception):
    """quakemigrate/util.py LUTPhasesException (raised at scan.py:624-629)."""


def marginal_window_indices(max_coa: np.ndarray, marginal_window: float, sampling_rate: int):
    """
    Sample range the reference keeps for the marginalised map.

    ``Event.add_compute_output`` sets ``otime`` to the time of the maximum of COA
    (io/event.py:239-240, first maximum); ``Event.trim2window`` keeps rows with
    ``otime - mw <= DT <= otime + mw`` and slices ``map4d[..., idx[0]:idx[-1]]`` -- END-EXCLUSIVE,
    i.e. the last kept row is dropped from the map (io/event.py:427-435).
    Returns ``(i0, i1, imax)`` with the map window ``[i0, i1)``.
    """
    imax = int(np.argmax(max_coa))
    # rows with |DT - otime| <= mw are kept (an inclusive comparison of times, io/event.py:427-431):
    # floor(mw * sr) samples per side, NOT round -- mw = 0.15 s at 50 Hz keeps 7, not 8
    half = int(np.floor(marginal_window * sampling_rate + 1e-9))
    first = max(0, imax - half)
    last = min(len(max_coa) - 1, imax + half)
    return first, last, imax


class QuakeScan:
    """
    GPU re-host of the compute core of ``quakemigrate.signal.scan.QuakeScan``.

    Parameters mirror the reference constructor where they touch the hot path (scan.py:204-280):
    ``lut``, ``onset``, and the attributes ``marginal_window``, ``timestep``, ``threads``
    (accepted, ignored).  Archive reading, output writing, picking and magnitudes are out of
    scope; callers hand in pre-processed waveform containers.
    """

    def __init__(self, lut, onset, engine: Engine | None = None, **kwargs):
        if not isinstance(onset, Onset) and not hasattr(onset, "calculate_onsets"):
            raise TypeError("onset must implement the Onset plugin interface")  # scan.py:209-212
        self.lut = lut
        self.onset = onset
        self.onset.post_pad = lut.max_traveltime  # scan.py:213
        self.engine = engine or getattr(onset, "engine", None) or Engine(0)
        if getattr(onset, "engine", "absent") is None:
            onset.engine = self.engine
        self.marginal_window = kwargs.get("marginal_window", 2.0)
        self.timestep = kwargs.get("timestep", 120.0)
        self.threads = kwargs.get("threads", 1)
        self.scan_rate = kwargs.get("scan_rate", self.onset.sampling_rate)
        self.pre_pad = self.post_pad = 0.0
        self._lut_cache: dict = {}
        # serve the int32 table on the device from resident float64 grids when the LUT can
        self.serve_on_device = kwargs.get("serve_on_device", hasattr(lut, "to_device"))
        self._grids = None

    # -- travel-time cache -----------------------------------------------------------------------
    def _device_lut(self, sampling_rate, availability):
        key = (int(sampling_rate), tuple(availability.items()))
        if key not in self._lut_cache:
            if len(self._lut_cache) >= 4:  # bound device memory: drop the oldest pattern
                old = next(iter(self._lut_cache))
                self._lut_cache.pop(old).free()
            try:
                if self.serve_on_device:
                    if self._grids is None:
                        self._grids = self.lut.to_device(self.engine)
                    grids, index = self._grids
                    rows = self.lut.rows_for(availability, index)
                    self._lut_cache[key] = self.engine.serve_lut(grids, rows, sampling_rate)
                else:
                    traveltimes = self.lut.serve_traveltimes(sampling_rate, availability)
                    self._lut_cache[key] = self.engine.upload_lut(traveltimes)
            except KeyError as e:
                raise LUTPhasesException(
                    f"Attempting to migrate phases; but traveltimes for {e} not found in the LUT."
                )
        return self._lut_cache[key]

    def set_pads(self, timespan: float):
        """``pre_pad, post_pad = onset.pad(timespan)`` (scan.py:425 detect, :494 locate)."""
        self.pre_pad, self.post_pad = self.onset.pad(timespan)
        return self.pre_pad, self.post_pad

    # -- the hot path ------------------------------------------------------------------------------
    def _prepare(self, data):
        onsets, onset_data = self.onset.calculate_onsets(data)
        dlut = self._device_lut(onset_data.sampling_rate, onset_data.availability)
        fsmp = time2sample(self.pre_pad, onset_data.sampling_rate)
        lsmp = time2sample(self.post_pad, onset_data.sampling_rate)
        avail = int(np.sum([value for _, value in onset_data.availability.items()]))
        dev = getattr(onset_data, "device_onsets", None)
        owned = False
        if not isinstance(dev, DeviceOnsets):
            n_onsets, t_samples = np.shape(onsets)
            if n_onsets != dlut.n_rows:  # lib.py:105-108
                raise ValueError(
                    f"Mismatch between number of stations for data and LUT, {n_onsets}:{dlut.n_rows}"
                )
            dev, owned = self.engine.upload_onsets(onsets), True
        return onsets, onset_data, dlut, dev, owned, fsmp, lsmp, avail

    def _compute(self, data, event=None):
        """
        Reference-shaped entry (scan.py:594-647).  Detect (``event is None``):
        ``(time, max_coa, max_coa_n, coord, onset_data)``.  Locate: additionally the FULL
        ``map4d`` (materialised through the marginal-window kernel over the whole span; large).
        """
        onsets, onset_data, dlut, dev, owned, fsmp, lsmp, avail = self._prepare(data)
        n_samples = dev.n_total - fsmp - lsmp
        if n_samples <= 0:  # lib.py:109-110
            raise ValueError("Data array smaller than coalescence array.")
        self.engine.detect_device(dlut, dev, fsmp, lsmp, avail)
        max_coa, max_coa_n, max_idx = self.engine.detect_fetch(n_samples)
        coord = self.lut.index2coord(max_idx, unravel=True)
        try:
            if event is None:
                time = data.starttime + self.pre_pad
                return time, max_coa, max_coa_n, coord, onset_data
            _, map4d = self.engine.locate_window(dlut, dev, fsmp, lsmp, avail, 0, n_samples,
                                                 return_map4d=True)
            times = data.starttime + self.pre_pad + np.arange(n_samples) / onset_data.sampling_rate
            return times, max_coa, max_coa_n, coord, map4d, onset_data
        finally:
            if owned:
                dev.free()

    def compute_locate(self, data, marginal_window: float | None = None, return_map4d: bool = False,
                       fit_location: bool = False, return_coa_map: bool = True):
        """
        Fast locate: the max-coalescence traces over the whole span, then only the marginal
        window around the peak (io/event.py:421-439) is re-migrated on the device.

        Returns a dict with ``times``, ``max_coa``, ``max_coa_n``, ``coord``, ``window`` =
        ``(i0, i1)``, ``coa_map`` = ``sum_t map4d[..., i0:i1]`` (scan.py:720; divide by its
        ``nanmax`` for the normalised map of scan.py:721), optionally ``map4d`` (the window slab),
        and ``onset_data``.  With ``fit_location`` (and no slab requested) the map stays on the
        device for the location fits; ``return_coa_map=False`` then skips its 8*N-byte copy to the
        host altogether.
        """
        mw = self.marginal_window if marginal_window is None else marginal_window
        onsets, onset_data, dlut, dev, owned, fsmp, lsmp, avail = self._prepare(data)
        try:
            n_samples = dev.n_total - fsmp - lsmp
            if n_samples <= 0:
                raise ValueError("Data array smaller than coalescence array.")
            self.engine.detect_device(dlut, dev, fsmp, lsmp, avail)
            max_coa, max_coa_n, max_idx = self.engine.detect_fetch(n_samples)
            i0, i1, imax = marginal_window_indices(max_coa, mw, onset_data.sampling_rate)
            out = {
                "times": data.starttime + self.pre_pad + np.arange(n_samples) / onset_data.sampling_rate,
                "max_coa": max_coa, "max_coa_n": max_coa_n,
                "coord": self.lut.index2coord(max_idx, unravel=True),
                "window": (i0, i1), "peak_sample": imax, "onset_data": onset_data,
            }
            if i1 <= i0:  # degenerate single-row window: the reference slices an empty map
                out["coa_map"] = np.zeros(dlut.shape[:-1])
                out["map4d"] = np.zeros(dlut.shape[:-1] + (0,)) if return_map4d else None
            elif fit_location and not return_map4d and hasattr(self.engine, "locate_window_device"):
                # the marginal map never leaves the device: _calculate_location (scan.py:696-733)
                # gets its windows and moments from the handle (a few KB over PCIe per event)
                from .locate import calculate_location
                dmap = self.engine.locate_window_device(dlut, dev, fsmp, lsmp, avail, i0, i1)
                try:
                    out["coa_map"] = dmap.fetch() if return_coa_map else None
                    out["map4d"] = None
                    out["location"] = calculate_location(dmap, self.lut, self.engine)
                finally:
                    dmap.free()
            else:
                res = self.engine.locate_window(dlut, dev, fsmp, lsmp, avail, i0, i1, return_map4d=return_map4d)
                out["coa_map"], out["map4d"] = (res if return_map4d else (res, None))
                if fit_location:
                    from .locate import calculate_location
                    out["location"] = calculate_location(out["coa_map"], self.lut, self.engine)
            return out
        finally:
            if owned:
                dev.free()

    def detect_chunks(self, chunks):
        """
        The loop of ``_continuous_compute`` (scan.py:434-462) over already-read waveform chunks:
        yields ``(time, max_coa, max_coa_n, coord)`` per chunk.  Chunks must have been read with
        ``pre_pad`` / ``post_pad`` from :meth:`set_pads` so that they are independent.
        """
        for i, data in enumerate(chunks):
            logging.debug("detect chunk %d", i)
            time, max_coa, max_coa_n, coord, _ = self._compute(data)
            yield time, max_coa, max_coa_n, coord

    def continuous_compute(self, starttime, chunks, run, continuous_write=False, device_output=True):
        """
        ``_continuous_compute`` (scan.py:406-467) over already-read waveform chunks, INCLUDING its
        output stage: every timestep is appended to a :class:`~quakemigrate_b200.scanmseed.ScanmSEED`
        (scanmseed.py:74-150) and day files are written under ``<run>/detect/scanmseed``.

        ``chunks[i]`` is the waveform container of timestep ``i`` (read with the pads of
        ``onset.pad(timestep)``) or ``None`` when the archive had nothing usable -- the reference
        stores such a step as zeros (``ScanmSEED.empty``, scan.py:454-461).  With
        ``device_output`` the COA / COA_N / Z channels are quantised on the device and only int32
        channels plus grid coordinates come back (``qmb_detect_fetch_quantised``); otherwise the
        float64 traces are fetched and quantised on the host like the reference.  Both give
        byte-identical files.  The station availability of every step goes to
        ``<run>/detect/availability`` (io/availability.py:123-152).  Returns
        ``(ScanmSEED, [availability dict or None per step])``.
        """
        from .scanmseed import ScanmSEED

        coalescence = ScanmSEED(run, continuous_write, self.scan_rate)
        self.pre_pad, self.post_pad = self.onset.pad(self.timestep)
        ucf = getattr(self.lut, "unit_conversion_factor", 1.0)
        availability = []
        for i, data in enumerate(chunks):
            if data is None:
                coalescence.empty(starttime, self.timestep, i, f"no data for timestep {i}", ucf)
                availability.append(None)
                continue
            try:
                if device_output:
                    onsets, onset_data, dlut, dev, owned, fsmp, lsmp, avail = self._prepare(data)
                    try:
                        n_samples = dev.n_total - fsmp - lsmp
                        if n_samples <= 0:
                            raise ValueError("Data array smaller than coalescence array.")
                        self.engine.detect_device(dlut, dev, fsmp, lsmp, avail)
                        coalescence.append_from_engine(data.starttime + self.pre_pad, self.engine,
                                                       n_samples, self.lut, ucf)
                    finally:
                        if owned:
                            dev.free()
                else:
                    time, max_coa, max_coa_n, coord, onset_data = self._compute(data)
                    coalescence.append(time, max_coa, max_coa_n, coord, ucf)
                availability.append(dict(onset_data.availability))
            except DataAvailabilityException as e:
                coalescence.empty(starttime, self.timestep, i, str(e), ucf)
                availability.append(None)
        if not coalescence.written:
            coalescence.write()
        if any(a is not None for a in availability):  # scan.py:467
            from .availability import write_availability
            from .scanmseed import to_us
            times = [to_us(starttime) + int(round(self.timestep * i * 1e6)) for i in range(len(availability))]
            write_availability(run, times, availability)
        return coalescence, availability

    def close(self):
        for dlut in self._lut_cache.values():
            dlut.free()
        self._lut_cache.clear()
        if self._grids is not None:
            self._grids[0].free()
            self._grids = None
