# -*- coding: utf-8 -*-
"""
TEST INFRASTRUCTURE ONLY: a CPU stand-in for ``quakemigrate_b200.core.Engine`` built from the
oracle (reference C / scipy), used by ``-m "not gpu"`` tests to exercise the HOST logic of
``QuakeScan`` / ``STALTAOnset`` / ``ScanmSEED`` (row ordering, pads, caches, file naming, error
paths) where no GPU exists.  It implements the documented semantics of the C ABI entry points the
host code calls -- it is not a fallback and nothing in the product can reach it.
"""
import numpy as np

from oracle import preprocess_oracle as PO
from oracle import qm_oracle as O
from quakemigrate_b200.core import DeviceOnsets


class _Lut:
    def __init__(self, tau):
        self.tau = np.ascontiguousarray(tau, dtype=np.int32)
        self.shape = self.tau.shape
        self.freed = False

    @property
    def n_rows(self):
        return self.shape[-1]

    def free(self):
        self.freed = True


class _Onsets(DeviceOnsets):
    def __init__(self, onsets):
        self.lin = np.ascontiguousarray(onsets, dtype=float)
        self.n_rows, self.n_total = self.lin.shape
        self._h = None

    def free(self):
        pass


class _Traces:
    def __init__(self, a):
        self.a = a
        self.n_channels, self.n = a.shape

    @property
    def shape(self):
        return self.a.shape

    def fetch(self, chan0=0, n_chan=None):
        n_chan = self.n_channels - chan0 if n_chan is None else n_chan
        return self.a[chan0:chan0 + n_chan].copy()

    def free(self):
        pass


class StandInEngine:
    def __init__(self):
        self.last = None
        self.uploads = 0

    # -- tables ------------------------------------------------------------------------------
    def upload_lut(self, traveltimes, node_offset=0, n_nodes_global=None):
        self.uploads += 1
        return _Lut(traveltimes)

    def upload_traveltime_grids(self, grids):
        return _Traces(np.ascontiguousarray(grids, dtype=float).reshape(len(grids), -1)), np.shape(grids)

    # -- onsets ------------------------------------------------------------------------------
    def upload_onsets(self, onsets):
        return _Onsets(onsets)

    def _onsets(self, wf, row_channels, nsta, nlta, position, transform, min_onset_value):
        fn = O.overlapping_sta_lta if position == "classic" else O.centred_sta_lta
        nsta = np.broadcast_to(np.asarray(nsta), (len(row_channels),))
        nlta = np.broadcast_to(np.asarray(nlta), (len(row_channels),))
        sig = {"energy": lambda x: x ** 2, "abs": np.abs, "env": PO.envelope,
               "env_squared": lambda x: PO.envelope(x) ** 2}[transform]
        rows = []
        for r, chans in enumerate(row_channels):
            ons = np.array([fn(sig(wf[c]), int(nsta[r]), int(nlta[r])) for c in chans])
            rows.append(np.clip(np.sqrt(np.sum(ons ** 2, axis=0) / len(ons)), min_onset_value, np.inf))
        return np.stack(rows)

    def onsets_from_waveforms(self, waveforms, row_channels, nsta, nlta, position="classic",
                              transform="energy", min_onset_value=0.4, return_onsets=False):
        lin = self._onsets(np.asarray(waveforms, dtype=float), row_channels, nsta, nlta, position, transform,
                           min_onset_value)
        return (_Onsets(lin), lin) if return_onsets else _Onsets(lin)

    def onsets_from_device_waveforms(self, traces, row_channels, nsta, nlta, **kw):
        return self.onsets_from_waveforms(traces.a, row_channels, nsta, nlta, **kw)

    # -- pre-processing ----------------------------------------------------------------------
    def upload_waveforms(self, waveforms):
        return _Traces(np.array(np.atleast_2d(waveforms), dtype=float))

    def alloc_waveforms(self, n_channels, n):
        return _Traces(np.zeros((n_channels, n)))

    def filter_waveforms(self, traces, sos=None, channels=None, detrend=True, taper_fraction=0.05,
                         zerophase=True, decimate=1, out=None, out_chan0=0, timing=False):
        from scipy.signal import detrend as sp_detrend
        from scipy.signal import sosfilt
        sel = range(traces.n_channels) if channels is None else channels
        if out is None:
            out = self.alloc_waveforms(len(sel) + out_chan0, (traces.n + decimate - 1) // decimate)
        for k, c in enumerate(sel):
            t = traces.a[c]
            if detrend:
                t = sp_detrend(sp_detrend(t, type="linear"), type="constant")
            if taper_fraction > 0:
                t = t * PO.cosine_taper_window(len(t), taper_fraction)
            if sos is not None:
                t = PO.sosfilt_zerophase(sos, t) if zerophase else sosfilt(sos, t)
            out.a[out_chan0 + k] = t[::decimate]
        return (out, 0.0) if timing else out

    def preprocess(self, waveforms, sos=None, detrend=True, taper_fraction=0.05, zerophase=True, decimate=1):
        raw = self.upload_waveforms(waveforms)
        return self.filter_waveforms(raw, sos, detrend=detrend, taper_fraction=taper_fraction,
                                     zerophase=zerophase, decimate=decimate).fetch()

    # -- detect ------------------------------------------------------------------------------
    def detect_device(self, lut, onsets, fsmp, lsmp, available):
        n_samples = onsets.n_total - fsmp - lsmp
        if lut.n_rows != onsets.n_rows:
            raise ValueError("row mismatch")
        self.last = O.detect_fused(onsets.lin, lut.tau, fsmp, lsmp, available, threads=2)
        assert len(self.last[0]) == n_samples
        return {"scan_ms": 0.0, "finalize_ms": 0.0, "total_ms": 0.0, "slow_samples": 0}

    def detect_fetch(self, n_samples):
        assert self.last is not None and len(self.last[0]) == n_samples
        return tuple(np.array(a) for a in self.last)

    def detect_fetch_quantised(self, n_samples, node_count, ll_corner, node_spacing, z_factor,
                               want=("coa", "coa_n", "ijk", "xyz", "z")):
        """include/qmb200.h, qmb_detect_fetch_quantised: the arithmetic of scanmseed.py:102-130."""
        max_coa, max_norm, idx = self.detect_fetch(n_samples)
        ijk = np.column_stack(np.unravel_index(idx, tuple(int(v) for v in node_count)))
        xyz = np.asarray(ll_corner, dtype=float) + ijk * np.asarray(node_spacing, dtype=float)
        out = {"coa": np.round(np.minimum(max_coa, 21474.0) * 1e5).astype(np.int32),
               "coa_n": np.round(np.minimum(max_norm, 21474.0) * 1e5).astype(np.int32),
               "ijk": ijk.astype(np.int32), "xyz": xyz,
               "z": np.round(xyz[:, 2] * z_factor).astype(np.int32)}
        return {k: v for k, v in out.items() if k in want}

    def locate_window(self, lut, onsets, fsmp, lsmp, available, win0, win1, return_map4d=False):
        m4 = O.migrate(onsets.lin, lut.tau, fsmp, lsmp, available, threads=2)[..., win0:win1]
        coa = m4.sum(axis=-1)
        return (coa, m4) if return_map4d else coa

    # -- locate, 3-D map stage (include/qmb200.h: qmb_gaufilt3d / qmb_map3d_peak / qmb_covfit3d) ---
    def gaufilt3d(self, map3d, sgm=0.8):
        return O.np_gaufilt3d(np.asarray(map3d, dtype=float), sgm)

    def map_peak(self, map3d):
        m = np.asarray(map3d, dtype=float)
        ijk = tuple(int(i) for i in np.unravel_index(np.nanargmax(m), m.shape))
        return ijk, float(np.nanmax(m)), float(np.nanmean(m))

    def covfit3d(self, map3d, node_spacing, thresh=0.90, win=None):
        assert not win, "windowed form is exercised on the GPU only"
        return O.np_covfit3d(np.asarray(map3d, dtype=float), np.broadcast_to(node_spacing, (3,)), thresh)
