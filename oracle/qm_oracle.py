# -*- coding: utf-8 -*-
"""
CPU oracle for the QuakeMigrate coalescence hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product
(``quakemigrate_b200`` + ``libqmb200.so``) never does; it fails loudly when its CUDA
library is missing instead of falling back to anything in here.

Three interchangeable back-ends compute the same functions:

``port``   ``oracle/_build/libqmoracle.so`` -- our plain-C restatement (``qm_oracle.c``),
           strict IEEE double.
``ref``    ``oracle/_ref/qmlib.so`` -- the UNMODIFIED reference C sources
           (``quakemigrate/core/src/{migratelib,onsetlib}.c``) compiled with the reference
           flags (``setup.py:111-119``).  Built by ``oracle/Makefile`` in the dev container;
           the built ``.so`` travels to the GPU box (no sources are copied).
``numpy``  pure numpy restatement below, for small cases and for reading.

The Python-level pre/post-processing that the reference performs around the C calls
(``quakemigrate/core/lib.py:93-110`` clip+log+shape checks, ``:154-158`` output allocation,
``:201-204`` STA/LTA header and pre-fill) is restated here with the same argument names so
that parity tests read like calls into ``quakemigrate.core``.

Parity status: PINNED -- against the reference KATs (``tests/test_onsets.py:23-39``), against
``oracle/_ref/qmlib.so`` on seeded random inputs, and against committed golden vectors
generated from it (``tests/golden/make_golden.py``).
"""

from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess

import numpy as np
import numpy.ctypeslib as clib

_HERE = pathlib.Path(__file__).resolve().parent

c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64
_dP = clib.ndpointer(dtype=np.double, flags="C_CONTIGUOUS")
_i32P = clib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64P = clib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

# numpy mirror of ``stalta_header`` (qmlib.h:34-38; lib.py:33-36)
stalta_header_t = np.dtype([("n", c_int32), ("nsta", c_int32), ("nlta", c_int32)], align=True)
_hdrP = clib.ndpointer(stalta_header_t, flags="C_CONTIGUOUS")

_MIGRATE_ARGS = [_dP, _i32P, _dP, c_int32, c_int32, c_int32, c_int32, c_int32, c_int64, c_int64]
_FINDMAX_ARGS = [_dP, _dP, _dP, _i64P, c_int32, c_int64, c_int64]
_STALTA_ARGS = [_dP, _hdrP, _dP]


def build(quiet: bool = True) -> None:
    """Run ``make -C oracle`` (builds the port; builds ``_ref`` only if /root/reference exists)."""
    subprocess.run(
        ["make", "-C", str(_HERE), "all"],
        check=True,
        stdout=subprocess.DEVNULL if quiet else None,
    )


class _Backend:
    """Uniform view over the port / reference shared objects."""

    def __init__(self, kind: str, path: pathlib.Path, names: dict):
        self.kind = kind
        self.path = path
        self.lib = ctypes.CDLL(str(path))
        self.migrate = getattr(self.lib, names["migrate"])
        self.migrate.argtypes, self.migrate.restype = _MIGRATE_ARGS, None
        self.find_max_coa = getattr(self.lib, names["find_max_coa"])
        self.find_max_coa.argtypes, self.find_max_coa.restype = _FINDMAX_ARGS, None
        for key in ("overlapping", "centred", "recursive"):
            fn = getattr(self.lib, names[key])
            fn.argtypes, fn.restype = _STALTA_ARGS, None
            setattr(self, f"{key}_sta_lta", fn)
        self.detect_fused = None
        if "detect_fused" in names:
            fn = getattr(self.lib, names["detect_fused"])
            fn.argtypes = [_dP, _i32P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int64,
                           _dP, _dP, _i64P, c_int64]
            fn.restype = None
            self.detect_fused = fn


_PORT_NAMES = dict(
    migrate="oracle_stack",
    find_max_coa="oracle_peak_trace",
    overlapping="oracle_stalta_overlapping",
    centred="oracle_stalta_centred",
    recursive="oracle_stalta_recursive",
    detect_fused="oracle_detect_fused",
)
_REF_NAMES = dict(
    migrate="migrate",
    find_max_coa="find_max_coa",
    overlapping="overlapping_sta_lta",
    centred="centred_sta_lta",
    recursive="recursive_sta_lta",
)

_cache: dict = {}


def have(kind: str) -> bool:
    return _path(kind).exists()


def _path(kind: str) -> pathlib.Path:
    return {
        "port": _HERE / "_build" / "libqmoracle.so",
        "ref": _HERE / "_ref" / "qmlib.so",
        "ref_strict": _HERE / "_ref" / "qmlib_strict.so",
    }[kind]


def backend(kind: str = "port") -> _Backend:
    """Load (building the port on demand) one of ``port`` / ``ref`` / ``ref_strict``."""
    if kind not in _cache:
        p = _path(kind)
        if not p.exists() and kind == "port":
            build()
        if not p.exists():
            raise FileNotFoundError(f"oracle backend '{kind}' not built: {p}")
        _cache[kind] = _Backend(kind, p, _PORT_NAMES if kind == "port" else _REF_NAMES)
    return _cache[kind]


def default_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()


# ----------------------------------------------------------------------------------------
# quakemigrate.core.lib look-alikes (C back-ends)
# ----------------------------------------------------------------------------------------
def log_onsets(onsets: np.ndarray) -> np.ndarray:
    """clip(0.01, inf) then natural log -- quakemigrate/core/lib.py:93-94."""
    return np.log(np.clip(onsets, 0.01, np.inf))


def migrate(onsets, traveltimes, first_idx, last_idx, available, threads=1, kind="port"):
    """``quakemigrate.core.lib.migrate`` (lib.py:53-125): returns map4d (nx, ny, nz, T) float64."""
    onsets = np.ascontiguousarray(log_onsets(onsets), dtype=np.double)
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
    backend(kind).migrate(
        onsets, np.ascontiguousarray(traveltimes, dtype=np.int32), map4d,
        c_int32(first_idx), c_int32(last_idx), c_int32(n_samples), c_int32(n_onsets),
        c_int32(int(available)), c_int64(n_nodes), c_int64(threads),
    )
    return map4d


def find_max_coa(map4d, threads=1, kind="port"):
    """``quakemigrate.core.lib.find_max_coa`` (lib.py:132-170)."""
    *grid_dimensions, n_samples = map4d.shape
    n_nodes = int(np.prod(grid_dimensions))
    max_coa = np.zeros(n_samples, dtype=np.double)
    max_norm_coa = np.zeros(n_samples, dtype=np.double)
    max_coa_idx = np.zeros(n_samples, dtype=np.int64)
    backend(kind).find_max_coa(
        map4d, max_coa, max_norm_coa, max_coa_idx,
        c_int32(n_samples), c_int64(n_nodes), c_int64(threads),
    )
    return max_coa, max_norm_coa, max_coa_idx


def _stalta(which, prefill, signal, nsta, nlta, kind):
    head = np.empty(1, dtype=stalta_header_t)
    head[:] = (len(signal), nsta, nlta)
    signal = np.ascontiguousarray(signal, dtype=np.double)
    onset = np.full(len(signal), prefill, dtype=np.double)
    getattr(backend(kind), which)(signal, head, onset)
    return onset


def overlapping_sta_lta(signal, nsta, nlta, kind="port"):
    """lib.py:176-208 (output pre-filled with 1.0)."""
    return _stalta("overlapping_sta_lta", 1.0, signal, nsta, nlta, kind)


def centred_sta_lta(signal, nsta, nlta, kind="port"):
    """lib.py:214-246 (output pre-filled with 1.0)."""
    return _stalta("centred_sta_lta", 1.0, signal, nsta, nlta, kind)


def recursive_sta_lta(signal, nsta, nlta, kind="port"):
    """lib.py:252-285 (output pre-filled with 0.0)."""
    return _stalta("recursive_sta_lta", 0.0, signal, nsta, nlta, kind)


def detect_fused(onsets, traveltimes, first_idx, last_idx, available, threads=1):
    """
    migrate + find_max_coa without materialising map4d (port back-end only); bit-identical to
    ``find_max_coa(migrate(...))`` of the port.  Used for full-size configs.
    """
    onsets = np.ascontiguousarray(log_onsets(onsets), dtype=np.double)
    n_luts = traveltimes.shape[-1]
    n_onsets, t_samples = onsets.shape
    n_samples = t_samples - first_idx - last_idx
    n_nodes = int(np.prod(traveltimes.shape[:-1]))
    if n_luts != n_onsets:
        raise ValueError(
            f"Mismatch between number of stations for data and LUT, {n_onsets}:{n_luts}"
        )
    max_coa = np.zeros(n_samples, dtype=np.double)
    max_norm_coa = np.zeros(n_samples, dtype=np.double)
    max_coa_idx = np.zeros(n_samples, dtype=np.int64)
    backend("port").detect_fused(
        onsets, np.ascontiguousarray(traveltimes, dtype=np.int32),
        c_int32(first_idx), c_int32(last_idx), c_int32(n_samples), c_int32(n_onsets),
        c_int32(int(available)), c_int64(n_nodes), max_coa, max_norm_coa, max_coa_idx,
        c_int64(threads),
    )
    return max_coa, max_norm_coa, max_coa_idx


# ----------------------------------------------------------------------------------------
# numpy restatement (small cases; readable definition of the path)
# ----------------------------------------------------------------------------------------
def np_migrate(onsets, traveltimes, first_idx, last_idx, available):
    """migratelib.c:40-65 in numpy: station-ascending accumulation, exp(sum / available)."""
    lo = log_onsets(np.asarray(onsets, dtype=np.double))
    *grid, n_rows = traveltimes.shape
    n_samples = lo.shape[1] - first_idx - last_idx
    tt = np.maximum(traveltimes.reshape(-1, n_rows), 0).astype(np.int64)
    t = np.arange(n_samples, dtype=np.int64)
    acc = np.zeros((tt.shape[0], n_samples), dtype=np.double)
    for s in range(n_rows):
        acc += lo[s][tt[:, s, None] + first_idx + t[None, :]]
    return np.exp(acc / float(available)).reshape(*grid, n_samples)


def np_find_max_coa(map4d):
    """migratelib.c:85-111 in numpy: first-index argmax, sequential node sum."""
    n_samples = map4d.shape[-1]
    flat = map4d.reshape(-1, n_samples)
    idx = np.argmax(flat, axis=0).astype(np.int64)  # numpy argmax returns the FIRST maximum
    mx = flat[idx, np.arange(n_samples)]
    total = flat[0].copy()
    for n in range(1, flat.shape[0]):  # same association as the serial C loop
        total += flat[n]
    return mx, mx * flat.shape[0] / total, idx


def np_overlapping_sta_lta(signal, nsta, nlta):
    """onsetlib.c:35-59 as literal window sums (no running-sum drift)."""
    x = np.asarray(signal, dtype=np.double)
    out = np.ones(len(x))
    for i in range(nlta - 1, len(x)):
        out[i] = x[i - nsta + 1:i + 1].sum() / x[i - nlta + 1:i + 1].sum() * (nlta / nsta)
    return out


def np_centred_sta_lta(signal, nsta, nlta):
    """onsetlib.c:79-108 as literal window sums."""
    x = np.asarray(signal, dtype=np.double)
    out = np.ones(len(x))
    for i in range(nlta - 1, len(x) - nsta):
        lta = x[i - nlta + 1:i + 1].sum()
        sta = x[i + 1:i + 1 + nsta].sum()
        out[i] = sta / lta * (nlta / nsta) if (lta > 0.0 or i == nlta - 1) else 1.0
    return out


def np_recursive_sta_lta(signal, nsta, nlta):
    """onsetlib.c:126-148."""
    x = np.asarray(signal, dtype=np.double)
    out = np.zeros(len(x))
    a, b = 1.0 / nsta, 1.0 / nlta
    sta = lta = 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(1, len(x)):
            sta = a * x[i] + (1 - a) * sta
            lta = b * x[i] + (1 - b) * lta
            out[i] = np.float64(sta) / np.float64(lta)
    if nlta < len(x):
        out[:nlta] = 1.0
    return out


def np_stalta_onset(channels, nsta, nlta, position="classic", transform="energy",
                    min_onset_value=0.4):
    """
    STALTAOnset._onset arithmetic (signal/onsets/stalta.py:516-548) for one station/phase:
    transform -> per-channel STA/LTA -> RMS over channels -> clip(min_onset_value).
    ``channels`` is (C, n).  Uses the port back-end for the STA/LTA.
    """
    ch = np.atleast_2d(np.asarray(channels, dtype=np.double))
    if transform == "energy":
        tr = ch ** 2
    elif transform == "abs":
        tr = np.abs(ch)
    else:
        raise ValueError(transform)
    fn = overlapping_sta_lta if position == "classic" else centred_sta_lta
    ons = np.array([fn(w, nsta, nlta) for w in tr])
    onset = np.sqrt(np.sum([o ** 2 for o in ons], axis=0) / len(ons))
    return np.clip(onset, min_onset_value, np.inf)


# ----------------------------------------------------------------------------------------
# locate: 3-D map stage -- numpy/scipy restatement, test use only
# ----------------------------------------------------------------------------------------
def np_gaussian_3d(nx, ny, nz, sgm):
    """``util.gaussian_3d`` (quakemigrate/util.py:76-116): samples at -(n-1)/2 .. (n-1)/2."""
    axes = [np.linspace(-(n - 1) / 2, (n - 1) / 2, n) for n in (nx, ny, nz)]
    gx, gy, gz = np.meshgrid(*axes, indexing="ij")
    s = np.repeat(sgm, 3) if np.isscalar(sgm) else np.asarray(sgm)
    return np.exp(-(gx * gx) / (2 * s[0] ** 2) - (gy * gy) / (2 * s[1] ** 2) - (gz * gz) / (2 * s[2] ** 2))


def np_gaufilt3d(map3d, sgm=0.8):
    """``QuakeScan._gaufilt3d`` (scan.py:1008-1043): smooth, mirror + normalise, twice."""
    from scipy.signal import fftconvolve

    flt = np_gaussian_3d(*map3d.shape, sgm)
    out = np.asarray(map3d, dtype=np.double)
    for _ in range(2):
        out = fftconvolve(out, flt, mode="same")
        out = out[::-1, ::-1, ::-1] / np.nanmax(out)
    return out


def np_covfit3d(coa_map, node_spacing, thresh=0.90):
    """Moments of ``QuakeScan._covfit3d`` (scan.py:966-1001): (expectation, covariance)."""
    w = np.where(coa_map > thresh, coa_map, np.nan)
    ssw = np.nansum(w)
    grids = np.meshgrid(*[np.arange(n) for n in coa_map.shape], indexing="ij")
    xs = [g * s for g, s in zip(grids, node_spacing)]
    e = [np.nansum(w * x) / ssw for x in xs]
    cov = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            cov[i, j] = np.nansum(w * (xs[i] - e[i]) * (xs[j] - e[j])) / ssw
    return np.array(e), cov
