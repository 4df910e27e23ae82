# -*- coding: utf-8 -*-
"""
Drop-in for :mod:`quakemigrate.core` backed by ``libqmb200.so`` (B200, sm_100a).

Tier 1 -- same five functions, same signatures, same exceptions as the reference bindings
(``quakemigrate/core/lib.py``): :func:`migrate`, :func:`find_max_coa`,
:func:`overlapping_sta_lta`, :func:`centred_sta_lta`, :func:`recursive_sta_lta`.  They call the
reference-compatible C symbols, which run on the GPU and return the full float64 volume through
host memory like the reference does.  ``migrate`` stacks float32 log-onsets in float32 and applies
a float64 ``exp``: its values agree with the reference to ~1e-6 relative (tests: 1e-5), not to the
last bit; ``find_max_coa`` on a given volume is bit-identical (same node order, same compares).
A failure inside the (void) C symbols raises :class:`QmbError` here.

Tier 2 -- :class:`Engine`: device-resident travel-time tables and onsets, fused
detect (no 4-D volume), marginal window for locate, node-sharded multi-GPU.

No CPU fallback exists: importing works anywhere, calling needs the CUDA library and a B200.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import QmbError, c_int32, c_int64, stalta_header_t  # noqa: F401

__all__ = ["migrate", "find_max_coa", "overlapping_sta_lta", "centred_sta_lta",
           "recursive_sta_lta", "Engine", "DeviceLUT", "DeviceOnsets", "DeviceGrids", "DeviceWaveforms", "DeviceMap",
           "QmbError"]


# ------------------------------------------------------------------------------------------------
# Tier 1: quakemigrate.core look-alikes
# ------------------------------------------------------------------------------------------------
def migrate(onsets, traveltimes, first_idx, last_idx, available, threads):
    """
    Computes 4-D coalescence map by migrating seismic phase onset functions.

    Same contract as ``quakemigrate.core.lib.migrate`` (lib.py:53-125): ``onsets``
    (nonsets, nsamples) float64 linear onsets, ``traveltimes`` (nx, ny, nz, nonsets) int32;
    returns ``map4d`` (nx, ny, nz, nsamples - first_idx - last_idx) float64.  ``threads`` is
    accepted for signature compatibility and ignored (the work runs on the GPU).

    Raises ``ValueError`` in the two cases the reference does (lib.py:105-110).
    """
    lib = _lib.load()
    # geometric mean as an arithmetic mean of logs; clip guards log(0) -- lib.py:93-94
    onsets = np.clip(onsets, 0.01, np.inf)
    onsets = np.ascontiguousarray(np.log(onsets), dtype=np.double)

    *grid_dimensions, n_luts = traveltimes.shape
    n_onsets, t_samples = onsets.shape
    n_samples = t_samples - first_idx - last_idx
    map4d = np.zeros(tuple(grid_dimensions) + (n_samples,), dtype=np.double)
    n_nodes = int(np.prod(grid_dimensions))

    if not n_luts == n_onsets:
        raise ValueError(
            f"Mismatch between number of stations for data and LUT, {n_onsets}:{n_luts}"
        )
    if onsets.size < n_samples + first_idx:
        raise ValueError("Data array smaller than coalescence array.")
    # The reference has no upper bound check on the travel times (migratelib.c:55-57): an offset
    # beyond the post-pad reads past the onset row.  Here that is an error, raised before any work.
    if traveltimes.size and int(np.max(traveltimes)) > last_idx:
        raise ValueError(
            f"Maximum travel time ({int(np.max(traveltimes))} samples) exceeds last_idx ({last_idx}): "
            "the post-pad must cover the largest travel time (scan.py:213)."
        )

    lib.migrate(
        onsets, np.ascontiguousarray(traveltimes, dtype=np.int32), map4d,
        c_int32(first_idx), c_int32(last_idx), c_int32(n_samples), c_int32(n_onsets),
        c_int32(int(available)), c_int64(n_nodes), c_int64(int(threads)),
    )
    _lib.check_legacy()  # the C symbol is void like the reference's: failures surface here
    return map4d


def find_max_coa(map4d, threads):
    """``quakemigrate.core.lib.find_max_coa`` (lib.py:132-170): (max_coa, max_norm_coa, max_coa_idx)."""
    lib = _lib.load()
    *grid_dimensions, n_samples = map4d.shape
    n_nodes = int(np.prod(grid_dimensions))
    max_coa = np.zeros(n_samples, dtype=np.double)
    max_norm_coa = np.zeros(n_samples, dtype=np.double)
    max_coa_idx = np.zeros(n_samples, dtype=np.int64)
    lib.find_max_coa(
        np.ascontiguousarray(map4d, dtype=np.double), max_coa, max_norm_coa, max_coa_idx,
        c_int32(n_samples), c_int64(n_nodes), c_int64(int(threads)),
    )
    _lib.check_legacy()
    return max_coa, max_norm_coa, max_coa_idx


def _legacy_stalta(name, prefill, signal, nsta, nlta):
    lib = _lib.load()
    if name != "recursive_sta_lta" and nsta > nlta:
        # onsetlib.c:44-58 subtracts samples it never added when the short window is the longer one
        raise ValueError(f"nsta ({nsta}) must not exceed nlta ({nlta})")
    head = np.empty(1, dtype=stalta_header_t)
    head[:] = (len(signal), nsta, nlta)
    signal = np.ascontiguousarray(signal, dtype=np.double)
    onset = np.full(len(signal), prefill, dtype=np.double)
    getattr(lib, name)(signal, head, onset)
    _lib.check_legacy()
    return onset


def overlapping_sta_lta(signal, nsta, nlta):
    """Classic STA/LTA, value at the window end (lib.py:176-208; onsetlib.c:35-59)."""
    return _legacy_stalta("overlapping_sta_lta", 1.0, signal, nsta, nlta)


def centred_sta_lta(signal, nsta, nlta):
    """Centred STA/LTA (lib.py:214-246; onsetlib.c:79-108)."""
    return _legacy_stalta("centred_sta_lta", 1.0, signal, nsta, nlta)


def recursive_sta_lta(signal, nsta, nlta):
    """Recursive STA/LTA (lib.py:252-285; onsetlib.c:126-148)."""
    return _legacy_stalta("recursive_sta_lta", 0.0, signal, nsta, nlta)


# ------------------------------------------------------------------------------------------------
# Tier 2: device-resident engine
# ------------------------------------------------------------------------------------------------
class DeviceLUT:
    """Handle to a travel-time table (node slab) resident on the device."""

    def __init__(self, engine, handle, shape, node_offset, n_nodes_global):
        self.engine, self._h = engine, handle
        self.shape = tuple(shape)  # (nx_local, ny, nz, n_rows)
        self.node_offset, self.n_nodes_global = node_offset, n_nodes_global

    @property
    def n_rows(self):
        return self.shape[-1]

    @property
    def n_nodes_local(self):
        return int(np.prod(self.shape[:-1]))

    def info(self) -> dict:
        inf = _lib.LutInfo()
        _lib.check(_lib.load().qmb_lut_info(self._h, ctypes.byref(inf)))
        return {name: getattr(inf, name) for name, _ in inf._fields_}

    def free(self):
        if self._h:
            _lib.load().qmb_lut_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceGrids:
    """Handle to float64 travel-time grids (seconds) resident on the device."""

    def __init__(self, engine, handle, shape):
        self.engine, self._h, self.shape = engine, handle, tuple(shape)

    def decimate(self, df) -> "DeviceGrids":
        """``LUT.decimate`` (lut.py:102-140) on the device: every ``df``-th node per axis, centred."""
        d = np.ascontiguousarray(np.broadcast_to(np.asarray(df, dtype=np.int32), (3,)))
        h = ctypes.c_void_p()
        _lib.check(_lib.load().qmb_ttgrids_decimate(self.engine._h, self._h, d, ctypes.byref(h)))
        shape = np.zeros(4, dtype=np.int32)
        _lib.check(_lib.load().qmb_ttgrids_shape(h, shape))
        return DeviceGrids(self.engine, h, tuple(int(v) for v in shape))

    def free(self):
        if self._h:
            _lib.load().qmb_ttgrids_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceOnsets:
    """Handle to log-onset rows resident on the device."""

    def __init__(self, engine, handle, n_rows, n_total):
        self.engine, self._h = engine, handle
        self.n_rows, self.n_total = n_rows, n_total

    def free(self):
        if self._h:
            _lib.load().qmb_onsets_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceMap:
    """
    Handle to a float64 ``(nx, ny, nz)`` map resident on the device: the marginalised coalescence
    map of locate() and its smoothed copy (scan.py:719-733).  Reductions and small windows come
    back to the host; the map itself only when :meth:`fetch` is called.
    """

    def __init__(self, engine, handle, shape):
        self.engine, self._h, self.shape = engine, handle, tuple(int(v) for v in shape)

    @property
    def nbytes(self):
        return 8 * int(np.prod(self.shape))

    def fetch(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.double)
        _lib.check(_lib.load().qmb_map3d_fetch(self._h, out))
        return out

    def normalise(self) -> "DeviceMap":
        """``map / nanmax(map)`` in place (scan.py:721)."""
        _lib.check(_lib.load().qmb_map3d_normalise(self.engine._h, self._h))
        return self

    def peak(self):
        """(nanargmax as ijk tuple, nanmax, nanmean) -- scan.py:869,878,972."""
        idx, mx, mean = c_int64(-1), ctypes.c_double(0.0), ctypes.c_double(0.0)
        _lib.check(_lib.load().qmb_map3d_peak_device(self.engine._h, self._h, ctypes.byref(idx),
                                                     ctypes.byref(mx), ctypes.byref(mean)))
        return tuple(int(i) for i in np.unravel_index(idx.value, self.shape)), mx.value, mean.value

    def gaufilt(self, sgm: float = 0.8) -> "DeviceMap":
        """``QuakeScan._gaufilt3d`` (scan.py:1008-1043): a new device map."""
        h = ctypes.c_void_p()
        _lib.check(_lib.load().qmb_map3d_gaufilt(self.engine._h, self._h, float(sgm), ctypes.byref(h)))
        return DeviceMap(self.engine, h, self.shape)

    def window(self, lo, hi) -> np.ndarray:
        """The box ``[lo, hi)`` as a host array (the 5^3 / 7^3 fit windows, scan.py:1045-1070)."""
        out = np.empty(tuple(int(h - l) for l, h in zip(lo, hi)), dtype=np.double)
        _lib.check(_lib.load().qmb_map3d_window(self.engine._h, self._h, int(lo[0]), int(hi[0]), int(lo[1]),
                                                int(hi[1]), int(lo[2]), int(hi[2]), out))
        return out

    def covfit(self, node_spacing, thresh: float = 0.90, win: int | None = None):
        """Reductions of ``QuakeScan._covfit3d`` (scan.py:966-1001): (expectation, covariance)."""
        sp = np.ascontiguousarray(np.broadcast_to(np.asarray(node_spacing, dtype=np.double), (3,)))
        out = np.empty(12, dtype=np.double)
        _lib.check(_lib.load().qmb_map3d_covfit(self.engine._h, self._h, sp, float(thresh), int(win or 0), out))
        return out[:3].copy(), out[3:].reshape(3, 3).copy()

    def free(self):
        if self._h:
            _lib.load().qmb_map3d_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceWaveforms:
    """Handle to a block of float64 traces ``(n_channels, n)`` resident on the device."""

    def __init__(self, engine, handle, n_channels, n):
        self.engine, self._h = engine, handle
        self.n_channels, self.n = int(n_channels), int(n)

    @property
    def shape(self):
        return (self.n_channels, self.n)

    def upsample(self, upfactor: int) -> "DeviceWaveforms":
        """The interpolation of ``util.upsample`` (util.py:544-551) on the device: a new block."""
        h = ctypes.c_void_p()
        _lib.check(_lib.load().qmb_waveforms_upsample(self.engine._h, self._h, int(upfactor), ctypes.byref(h)))
        return DeviceWaveforms(self.engine, h, self.n_channels, (self.n - 1) * int(upfactor) + 1)

    def fetch(self, chan0: int = 0, n_chan: int | None = None) -> np.ndarray:
        n_chan = self.n_channels - chan0 if n_chan is None else n_chan
        out = np.empty((n_chan, self.n), dtype=np.double)
        _lib.check(_lib.load().qmb_waveforms_fetch(self._h, int(chan0), int(n_chan), out))
        return out

    def free(self):
        if self._h:
            _lib.load().qmb_waveforms_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One GPU context: streams, scratch, optional NCCL communicator."""

    def __init__(self, device: int = 0, variant: int | None = None, precision: str | None = None,
                 qgroup: int | None = None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_ctx_create(int(device), ctypes.byref(h)))
        self._h = h
        self.device = device
        self.world, self.rank = 1, 0
        if variant is not None:
            # kernel tiling variant (consumer warps x nodes per warp x samples per lane), see
            # QMB_VARIANTS in csrc/qmb_internal.cuh; applies to tables uploaded afterwards
            _lib.check(self._lib.qmb_ctx_set_variant(self._h, int(variant)))
        if precision is not None:
            self.set_scan_precision(precision)
        if qgroup is not None:
            # rows per carry-free group of the fixed-point scan: 1 (16-bit rows), 2, 4, 8 (13-bit)
            _lib.check(self._lib.qmb_ctx_set_scan_qgroup(self._h, int(qgroup)))

    # -- lifetime ---------------------------------------------------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qmb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._lib.qmb_ctx_sync(self._h))

    def set_scan_precision(self, precision: str):
        """
        Operand format of the detect scan's gathers: ``"auto"`` (default: 16-bit fixed-point rows,
        two samples per shared-memory word, whenever ``max_norm_coa`` stays clearly inside the
        1e-5 bar; float32 rows on tiny grids), ``"q16"`` or ``"fp32"`` to force one.  Arg-max
        indices and ``max_coa`` are decided in float64 among the candidates inside the error band
        either way.
        """
        modes = {"auto": 0, "fp32": 1, "q16": 2}
        if precision not in modes:
            raise ValueError("precision must be 'auto', 'q16' or 'fp32'")
        _lib.check(self._lib.qmb_ctx_set_scan_precision(self._h, modes[precision]))

    @property
    def last_scan_bytes_per_gather(self) -> int:
        n = ctypes.c_int(0)
        _lib.check(self._lib.qmb_ctx_last_scan_bytes_per_gather(self._h, ctypes.byref(n)))
        return n.value

    @property
    def last_scan_info(self) -> dict:
        """
        What the last detect scan did (``qmb_scan_info_t``): operand format, the node-sum
        verification of the 16-bit scan (windows, samples, worst relative difference, whether the
        scan was redone with float32 rows), candidate / slow-path counts and CUDA-event times.
        """
        inf = _lib.ScanInfo()
        _lib.check(self._lib.qmb_ctx_last_scan_info(self._h, ctypes.byref(inf)))
        return {name: getattr(inf, name) for name, _ in inf._fields_}

    @property
    def launch_count(self) -> int:
        n = c_int64(0)
        _lib.check(self._lib.qmb_ctx_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def smem_pipe_peak(self, iters: int = 20000) -> float:
        """Measured ceiling of the shared-memory pipe: LDS.32 + FADD gather-adds per second."""
        out = ctypes.c_double(0.0)
        _lib.check(self._lib.qmb_microbench_smem_pipe(self._h, int(iters), ctypes.byref(out)))
        return out.value

    def timer_start(self):
        """Record a CUDA event on the engine's stream (see :meth:`timer_stop`)."""
        _lib.check(self._lib.qmb_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        """Device milliseconds since :meth:`timer_start` (CUDA events on the launching stream)."""
        ms = ctypes.c_float(0.0)
        _lib.check(self._lib.qmb_ctx_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def flush_l2(self):
        """Evict L2 (benchmark hygiene between timed iterations)."""
        _lib.check(self._lib.qmb_ctx_flush_l2(self._h))

    @staticmethod
    def pin(array: np.ndarray):
        """Page-lock a host array in place (cudaHostRegister) so copies run at PCIe rate."""
        _lib.check(_lib.load().qmb_host_register(array.ctypes.data_as(ctypes.c_void_p), array.nbytes))

    @staticmethod
    def unpin(array: np.ndarray):
        _lib.check(_lib.load().qmb_host_unregister(array.ctypes.data_as(ctypes.c_void_p)))

    # -- multi-GPU --------------------------------------------------------------------------
    @staticmethod
    def new_comm_id() -> bytes:
        buf = ctypes.create_string_buffer(_lib.QMB_NCCL_ID_BYTES)
        _lib.check(_lib.load().qmb_comm_unique_id(buf))
        return buf.raw

    def init_comm(self, comm_id: bytes, world: int, rank: int):
        _lib.check(self._lib.qmb_comm_init(self._h, comm_id, int(world), int(rank)))
        self.world, self.rank = world, rank

    # -- inputs -----------------------------------------------------------------------------
    def upload_lut(self, traveltimes: np.ndarray, node_offset: int = 0,
                   n_nodes_global: int | None = None) -> DeviceLUT:
        """``traveltimes``: int32 (nx_local, ny, nz, n_rows) as ``LUT.serve_traveltimes`` returns."""
        if traveltimes.ndim != 4:
            raise ValueError("traveltimes must have shape (nx, ny, nz, n_onsets)")
        tt = np.ascontiguousarray(traveltimes, dtype=np.int32)
        nx, ny, nz, s = tt.shape
        n_glob = int(n_nodes_global) if n_nodes_global else nx * ny * nz
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_lut_upload(self._h, tt, nx, ny, nz, s, int(node_offset), n_glob,
                                            ctypes.byref(h)))
        return DeviceLUT(self, h, tt.shape, int(node_offset), n_glob)

    def upload_lut_range(self, rows: np.ndarray, ny: int, nz: int, node_first: int,
                         n_nodes_global: int) -> DeviceLUT:
        """
        A slab given as a contiguous range of the C-order node list: ``rows`` is int32
        ``(n_nodes_local, n_rows)`` = ``traveltimes.reshape(-1, n_rows)[node_first : node_first +
        n_nodes_local]``; both ends must fall on whole z-columns (multiples of ``nz``) but may lie
        inside an x-plane -- see :func:`quakemigrate_b200.sharding.balanced_range`.
        """
        tt = np.ascontiguousarray(rows, dtype=np.int32)
        if tt.ndim != 2:
            raise ValueError("rows must have shape (n_nodes_local, n_onsets)")
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_lut_upload_range(self._h, tt, int(ny), int(nz), tt.shape[1], int(node_first),
                                                  tt.shape[0], int(n_nodes_global), ctypes.byref(h)))
        return DeviceLUT(self, h, tt.shape, int(node_first), int(n_nodes_global))

    def upload_traveltime_grids(self, grids: np.ndarray) -> "DeviceGrids":
        """``grids``: float64 (n_grids, nx, ny, nz) travel times in seconds, kept on the device."""
        g = np.ascontiguousarray(grids, dtype=np.double)
        if g.ndim != 4:
            raise ValueError("grids must have shape (n_grids, nx, ny, nz)")
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_ttgrids_upload(self._h, g, *g.shape, ctypes.byref(h)))
        return DeviceGrids(self, h, g.shape)

    def serve_lut(self, grids: "DeviceGrids", rows, sampling_rate, x_range=None) -> DeviceLUT:
        """
        ``LUT.serve_traveltimes`` on the device: onset row r uses grid ``rows[r]``;
        tau = rint(tt * sampling_rate) as int32 (lut.py:538).  ``x_range`` selects a node slab.
        """
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        n_grids, nx, ny, nz = grids.shape
        x0, x1 = (0, nx) if x_range is None else x_range
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_lut_serve(self._h, grids._h, rows, len(rows), float(sampling_rate),
                                           int(x0), int(x1), ctypes.byref(h)))
        return DeviceLUT(self, h, (x1 - x0, ny, nz, len(rows)), x0 * ny * nz, nx * ny * nz)

    def upload_onsets(self, onsets: np.ndarray) -> DeviceOnsets:
        """``onsets``: float64 (n_rows, n_total) LINEAR onsets (clip + log happen on the device)."""
        on = np.ascontiguousarray(onsets, dtype=np.double)
        if on.ndim != 2:
            raise ValueError("onsets must have shape (n_onsets, n_samples)")
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_onsets_upload(self._h, on, on.shape[0], on.shape[1], ctypes.byref(h)))
        return DeviceOnsets(self, h, on.shape[0], on.shape[1])

    def onsets_from_waveforms(self, waveforms, row_channels, nsta, nlta, position="classic",
                              transform="energy", min_onset_value=0.4, return_onsets=False):
        """
        ``STALTAOnset._onset`` on the device for every row (stalta.py:516-548).

        waveforms: float64 (n_channels, n_total); row_channels: list (per onset row) of channel
        index lists; nsta / nlta: ints or per-row sequences (samples).
        Returns DeviceOnsets, or (DeviceOnsets, linear onsets ndarray) when ``return_onsets``.
        """
        wf = np.ascontiguousarray(waveforms, dtype=np.double)
        n_rows = len(row_channels)
        row_ptr = np.zeros(n_rows + 1, dtype=np.int32)
        row_ptr[1:] = np.cumsum([len(c) for c in row_channels])
        row_chan = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.int32) for c in row_channels]))
        nsta_a = np.ascontiguousarray(np.broadcast_to(np.asarray(nsta, dtype=np.int32), (n_rows,)))
        nlta_a = np.ascontiguousarray(np.broadcast_to(np.asarray(nlta, dtype=np.int32), (n_rows,)))
        out = np.empty((n_rows, wf.shape[1]), dtype=np.double) if return_onsets else None
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_onsets_from_waveforms(
            self._h, wf, wf.shape[0], wf.shape[1], n_rows, row_ptr, row_chan, nsta_a, nlta_a,
            _lib.STALTA_POSITION[position], _lib.TRANSFORM[transform], float(min_onset_value),
            out.ctypes.data_as(ctypes.c_void_p) if out is not None else None, ctypes.byref(h)))
        dev = DeviceOnsets(self, h, n_rows, wf.shape[1])
        return (dev, out) if return_onsets else dev

    # -- waveform pre-processing (stalta.py:137-211, util.py:479-520) ---------------------------
    def upload_waveforms(self, waveforms) -> DeviceWaveforms:
        """``waveforms``: float64 (n_channels, n) raw traces, copied to the device."""
        wf = np.ascontiguousarray(np.atleast_2d(waveforms), dtype=np.double)
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_waveforms_upload(self._h, wf, wf.shape[0], wf.shape[1], ctypes.byref(h)))
        return DeviceWaveforms(self, h, *wf.shape)

    def alloc_waveforms(self, n_channels: int, n: int) -> DeviceWaveforms:
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_waveforms_alloc(self._h, int(n_channels), int(n), ctypes.byref(h)))
        return DeviceWaveforms(self, h, n_channels, n)

    def filter_waveforms(self, traces: DeviceWaveforms, sos=None, channels=None, detrend=True,
                         taper_fraction=0.05, zerophase=True, decimate=1, out=None, out_chan0=0,
                         timing=False):
        """
        detrend (linear + constant) -> cosine taper -> IIR second-order sections (scipy layout,
        ``(n_sections, 6)``; ``None``: no filter) forward or zero-phase -> every ``decimate``-th
        sample, for the selected ``channels`` of ``traces`` (default all), written to channels
        ``out_chan0 ...`` of ``out`` (allocated when ``None``).  Returns ``out`` (and the device
        time of the kernels in ms when ``timing``).
        """
        sel = None if channels is None else np.ascontiguousarray(channels, dtype=np.int32)
        n_sel = traces.n_channels if sel is None else len(sel)
        n_out = (traces.n + decimate - 1) // decimate
        if out is None:
            out = self.alloc_waveforms(n_sel + out_chan0, n_out)
        sos_a = None if sos is None else np.ascontiguousarray(np.atleast_2d(sos), dtype=np.double)
        if sos_a is not None and sos_a.shape[1] != 6:
            raise ValueError("sos must have shape (n_sections, 6)")
        ms = ctypes.c_float(0.0)
        _lib.check(self._lib.qmb_waveforms_filter(
            self._h, traces._h, None if sel is None else sel.ctypes.data_as(ctypes.c_void_p), int(n_sel),
            None if sos_a is None else sos_a.ctypes.data_as(ctypes.c_void_p),
            0 if sos_a is None else sos_a.shape[0], int(bool(detrend)), float(taper_fraction),
            int(bool(zerophase)), int(decimate), out._h, int(out_chan0),
            ctypes.cast(ctypes.byref(ms), ctypes.c_void_p) if timing else None))
        return (out, ms.value) if timing else out

    def preprocess(self, waveforms, sos=None, detrend=True, taper_fraction=0.05, zerophase=True,
                   decimate=1) -> np.ndarray:
        """One call, host arrays in and out (``qmb_preprocess``)."""
        wf = np.ascontiguousarray(np.atleast_2d(waveforms), dtype=np.double)
        sos_a = None if sos is None else np.ascontiguousarray(np.atleast_2d(sos), dtype=np.double)
        out = np.empty((wf.shape[0], (wf.shape[1] + decimate - 1) // decimate), dtype=np.double)
        _lib.check(self._lib.qmb_preprocess(
            self._h, wf, wf.shape[0], wf.shape[1],
            None if sos_a is None else sos_a.ctypes.data_as(ctypes.c_void_p),
            0 if sos_a is None else sos_a.shape[0], int(bool(detrend)), float(taper_fraction),
            int(bool(zerophase)), int(decimate), out))
        return out

    def onsets_from_device_waveforms(self, traces: DeviceWaveforms, row_channels, nsta, nlta,
                                     position="classic", transform="energy", min_onset_value=0.4,
                                     return_onsets=False):
        """:meth:`onsets_from_waveforms` on traces that already live on the device."""
        n_rows = len(row_channels)
        row_ptr = np.zeros(n_rows + 1, dtype=np.int32)
        row_ptr[1:] = np.cumsum([len(c) for c in row_channels])
        row_chan = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.int32) for c in row_channels]))
        nsta_a = np.ascontiguousarray(np.broadcast_to(np.asarray(nsta, dtype=np.int32), (n_rows,)))
        nlta_a = np.ascontiguousarray(np.broadcast_to(np.asarray(nlta, dtype=np.int32), (n_rows,)))
        out = np.empty((n_rows, traces.n), dtype=np.double) if return_onsets else None
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_onsets_from_device_waveforms(
            self._h, traces._h, n_rows, row_ptr, row_chan, nsta_a, nlta_a,
            _lib.STALTA_POSITION[position], _lib.TRANSFORM[transform], float(min_onset_value),
            out.ctypes.data_as(ctypes.c_void_p) if out is not None else None, ctypes.byref(h)))
        dev = DeviceOnsets(self, h, n_rows, traces.n)
        return (dev, out) if return_onsets else dev

    def envelope(self, signals) -> np.ndarray:
        """``np.abs(scipy.signal.hilbert(signals, axis=-1))`` (stalta.py:520-523) on the device."""
        sig = np.ascontiguousarray(np.atleast_2d(signals), dtype=np.double)
        out = np.empty_like(sig)
        _lib.check(self._lib.qmb_envelope(self._h, sig, sig.shape[0], sig.shape[1], out))
        return out

    def sta_lta(self, signals, nsta, nlta, position="classic", prefill=None):
        """Batched STA/LTA of (n_channels, n) signals; pre-fill follows lib.py:204,242,281."""
        sig = np.ascontiguousarray(np.atleast_2d(signals), dtype=np.double)
        if prefill is None:
            prefill = 0.0 if position == "recursive" else 1.0
        out = np.full(sig.shape, prefill, dtype=np.double)
        _lib.check(self._lib.qmb_sta_lta(self._h, sig, sig.shape[0], sig.shape[1], int(nsta), int(nlta),
                                         _lib.STALTA_POSITION[position], out))
        return out

    # -- detect -----------------------------------------------------------------------------
    def detect_device(self, lut: DeviceLUT, onsets: DeviceOnsets, fsmp: int, lsmp: int, available: int):
        """Run the fused scan on device-resident inputs; returns kernel timings (ms) as a dict."""
        ms = (ctypes.c_float * 4)()
        _lib.check(self._lib.qmb_detect_device(self._h, lut._h, onsets._h, int(fsmp), int(lsmp),
                                               int(available), ctypes.cast(ms, ctypes.c_void_p)))
        return {"scan_ms": ms[0], "finalize_ms": ms[1], "total_ms": ms[2], "slow_samples": int(ms[3])}

    def detect_fetch(self, n_samples: int):
        max_coa = np.empty(n_samples, dtype=np.double)
        max_norm = np.empty(n_samples, dtype=np.double)
        idx = np.empty(n_samples, dtype=np.int64)
        _lib.check(self._lib.qmb_detect_fetch(self._h, max_coa, max_norm, idx, int(n_samples)))
        return max_coa, max_norm, idx

    def detect_fetch_quantised(self, n_samples: int, node_count, ll_corner, node_spacing,
                               z_factor: float, want=("coa", "coa_n", "ijk", "xyz", "z")):
        """
        The last detect result in the integer form ``ScanmSEED.append`` stores
        (io/scanmseed.py:102-130), quantised on the device: dict with ``coa``, ``coa_n`` (int32,
        clip 21474, x1e5, half-to-even), ``ijk`` (int32 (n, 3), C-order unravel, lut.py:165-166),
        ``xyz`` (float64 (n, 3) grid coordinates, lut.py:177) and ``z`` (int32, ``xyz[:, 2]`` x
        ``z_factor`` = 1e3 * ucf).
        """
        n = int(n_samples)
        bufs = {"coa": np.empty(n, np.int32), "coa_n": np.empty(n, np.int32),
                "ijk": np.empty((n, 3), np.int32), "xyz": np.empty((n, 3), np.double),
                "z": np.empty(n, np.int32)}
        ptr = {k: (v.ctypes.data_as(ctypes.c_void_p) if k in want else None) for k, v in bufs.items()}
        _lib.check(self._lib.qmb_detect_fetch_quantised(
            self._h, np.ascontiguousarray(node_count, dtype=np.int32),
            np.ascontiguousarray(ll_corner, dtype=np.double),
            np.ascontiguousarray(node_spacing, dtype=np.double), float(z_factor),
            ptr["coa"], ptr["coa_n"], ptr["ijk"], ptr["xyz"], ptr["z"], n))
        return {k: v for k, v in bufs.items() if k in want}

    def detect(self, lut: DeviceLUT, onsets, fsmp: int, lsmp: int, available: int):
        """
        ``find_max_coa(migrate(onsets, lut, fsmp, lsmp, available))`` fused on the device
        (scan.py:635-638).  ``onsets`` may be a host array (linear, float64) or DeviceOnsets.
        """
        if isinstance(onsets, DeviceOnsets):
            self.detect_device(lut, onsets, fsmp, lsmp, available)
            return self.detect_fetch(onsets.n_total - fsmp - lsmp)
        on = np.ascontiguousarray(onsets, dtype=np.double)
        n_rows, n_total = on.shape
        if n_rows != lut.n_rows:
            raise ValueError(
                f"Mismatch between number of stations for data and LUT, {n_rows}:{lut.n_rows}"
            )
        n_samples = n_total - fsmp - lsmp
        if n_samples <= 0:
            raise ValueError("Data array smaller than coalescence array.")
        max_coa = np.empty(n_samples, dtype=np.double)
        max_norm = np.empty(n_samples, dtype=np.double)
        idx = np.empty(n_samples, dtype=np.int64)
        _lib.check(self._lib.qmb_detect(self._h, lut._h, on, n_rows, n_total, int(fsmp), int(lsmp),
                                        int(available), max_coa, max_norm, idx))
        return max_coa, max_norm, idx

    # -- locate: 3-D map stage ---------------------------------------------------------------
    def gaufilt3d(self, map3d: np.ndarray, sgm: float = 0.8) -> np.ndarray:
        """``QuakeScan._gaufilt3d`` (scan.py:1008-1043) on the device."""
        m = np.ascontiguousarray(map3d, dtype=np.double)
        if m.ndim != 3:
            raise ValueError("map3d must be 3-D")
        out = np.empty_like(m)
        _lib.check(self._lib.qmb_gaufilt3d(self._h, m, *m.shape, float(sgm), out))
        return out

    def map_peak(self, map3d: np.ndarray):
        """(nanargmax as ijk tuple, nanmax, nanmean) of a map (scan.py:869,878,972)."""
        m = np.ascontiguousarray(map3d, dtype=np.double)
        idx, mx, mean = c_int64(-1), ctypes.c_double(0.0), ctypes.c_double(0.0)
        _lib.check(self._lib.qmb_map3d_peak(self._h, m.reshape(-1), m.size, ctypes.byref(idx),
                                            ctypes.byref(mx), ctypes.byref(mean)))
        return tuple(int(i) for i in np.unravel_index(idx.value, m.shape)), mx.value, mean.value

    def covfit3d(self, map3d: np.ndarray, node_spacing, thresh: float = 0.90, win: int | None = None):
        """
        Weighted expectation and covariance of the nodes above ``thresh``
        (``QuakeScan._covfit3d`` reductions, scan.py:966-1001): returns
        ``(expectation (3,) in index*spacing units, covariance (3, 3))``.
        """
        m = np.ascontiguousarray(map3d, dtype=np.double)
        sp = np.ascontiguousarray(np.broadcast_to(np.asarray(node_spacing, dtype=np.double), (3,)))
        out = np.empty(12, dtype=np.double)
        _lib.check(self._lib.qmb_covfit3d(self._h, m, *m.shape, sp, float(thresh), int(win or 0), out))
        return out[:3].copy(), out[3:].reshape(3, 3).copy()

    def rbf_cubic_peak(self, weights, n: int, upscale: int = 10):
        """
        Peak of the cubic radial-basis interpolant on the ``upscale``-times finer window grid
        (``QuakeScan._splineloc``, scan.py:789-809): ``weights`` are the ``n**3`` weights of scipy's
        ``Rbf(..., function="cubic")`` for the window; returns (index triple into the fine grid as
        ``np.unravel_index(np.nanargmax(coa_map_int), ...)`` gives it, value).
        """
        w = np.ascontiguousarray(weights, dtype=np.double).reshape(-1)
        if w.size != n ** 3:
            raise ValueError("weights must hold n**3 values")
        idx, mx = c_int64(-1), ctypes.c_double(0.0)
        _lib.check(self._lib.qmb_rbf_cubic_peak(self._h, w, int(n), int(upscale), ctypes.byref(idx), ctypes.byref(mx)))
        m = (n - 1) * upscale + 1
        return tuple(int(i) for i in np.unravel_index(idx.value, (m, m, m))), mx.value

    def upload_map(self, map3d: np.ndarray) -> DeviceMap:
        m = np.ascontiguousarray(map3d, dtype=np.double)
        if m.ndim != 3:
            raise ValueError("map3d must be 3-D")
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_map3d_upload(self._h, m, *m.shape, ctypes.byref(h)))
        return DeviceMap(self, h, m.shape)

    # -- locate -----------------------------------------------------------------------------
    def locate_window_device(self, lut: DeviceLUT, onsets: DeviceOnsets, fsmp: int, lsmp: int,
                             available: int, win0: int, win1: int) -> DeviceMap:
        """
        :meth:`locate_window` whose result stays on the device: the un-normalised marginal map
        ``sum_t map4d[..., win0:win1]`` (scan.py:720) as a :class:`DeviceMap`.
        """
        h = ctypes.c_void_p()
        _lib.check(self._lib.qmb_locate_window_device(
            self._h, lut._h, onsets._h, int(fsmp), int(lsmp), int(available), int(win0), int(win1),
            ctypes.byref(h)))
        return DeviceMap(self, h, lut.shape[:-1])

    def locate_window(self, lut: DeviceLUT, onsets: DeviceOnsets, fsmp: int, lsmp: int, available: int,
                      win0: int, win1: int, return_map4d: bool = False):
        """
        Marginal window [win0, win1) (end-exclusive, io/event.py:433-435) of the 4-D map:
        returns ``coa_map`` (nx_local, ny, nz) = sum over the window (scan.py:720) and optionally
        the window slab (nx_local, ny, nz, win1 - win0).
        """
        grid = lut.shape[:-1]
        coa = np.empty(grid, dtype=np.double)
        slab = np.empty(grid + (win1 - win0,), dtype=np.double) if return_map4d else None
        _lib.check(self._lib.qmb_locate_window(
            self._h, lut._h, onsets._h, int(fsmp), int(lsmp), int(available), int(win0), int(win1),
            coa.ctypes.data_as(ctypes.c_void_p),
            slab.ctypes.data_as(ctypes.c_void_p) if slab is not None else None))
        return (coa, slab) if return_map4d else coa
