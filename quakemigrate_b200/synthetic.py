# -*- coding: utf-8 -*-
"""
Deterministic synthetic inputs for the coalescence hot path.

Everything here is host-side numpy and independent of both the CUDA library and the oracle: it
only manufactures the two arrays the hot path consumes -- travel-time look-up tables in the
layout ``LUT.serve_traveltimes`` returns (``quakemigrate/lut/lut.py:502-538``: int32,
``(nx, ny, nz, n_onsets)``, C-order) and multi-channel waveforms from which onset functions are
computed -- plus the pad arithmetic of ``Onset.pad`` / ``STALTAOnset.pre_pad`` / ``post_pad``
(``quakemigrate/signal/onsets/base.py:64-93``, ``stalta.py:600-635``, ``util.py:152-172,248-269``).

Travel times use the homogeneous-velocity formula of the reference's LUT builder
(``quakemigrate/lut/create_lut.py:262-265``: straight-ray distance / velocity).
"""

from __future__ import annotations

import dataclasses

import numpy as np

from .util import time2sample, trim2sample  # noqa: F401  (re-exported for the tests)


# ------------------------------------------------------------------------------------------
# pad -> sample arithmetic (host logic mirrored from the reference)
# ------------------------------------------------------------------------------------------
def stalta_pads(sta: float, lta: float, max_traveltime: float, timespan: float,
                sampling_rate: int) -> tuple[float, float]:
    """
    ``onset.pad(timespan)`` for a STALTAOnset whose ``post_pad`` was set from the LUT
    (scan.py:213): pre_pad = lta + 3 sta (stalta.py:600-608), post_pad = ceil(ttmax + 2 lta)
    (stalta.py:626-635), each widened by ceil(6 % of the padded span) (base.py:84-93).
    """
    pre = lta + 3.0 * sta
    post = float(np.ceil(max_traveltime + 2.0 * lta))
    span = timespan + pre + post
    pre_pad = trim2sample(pre + np.ceil(span * 0.06), sampling_rate)
    post_pad = trim2sample(post + np.ceil(span * 0.06), sampling_rate)
    return pre_pad, post_pad


# ------------------------------------------------------------------------------------------
# configurations (BASELINE.json "configs")
# ------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class SynthConfig:
    name: str
    grid: tuple[int, int, int]
    n_rows: int                 # S: onset rows handed to migrate (station x phase pairs)
    n_samples: int              # T: output samples per call
    position: str = "classic"   # STA/LTA flavour
    sampling_rate: int = 100
    node_spacing_km: float = 0.5
    vp: float = 5.0
    vs: float = 5.0 / 1.73
    sta: float = 0.2
    lta: float = 1.0
    seed: int = 0
    n_events: int = 8

    @property
    def n_nodes(self) -> int:
        return int(np.prod(self.grid))

    def scaled(self, grid=None, n_samples=None, n_rows=None, name=None) -> "SynthConfig":
        return dataclasses.replace(
            self,
            grid=tuple(grid) if grid is not None else self.grid,
            n_samples=n_samples if n_samples is not None else self.n_samples,
            n_rows=n_rows if n_rows is not None else self.n_rows,
            name=name or (self.name + "-scaled"),
        )


CONFIGS = {
    # cfg1 is the bundled Iceland example geometry (real LUT fixture when available); the
    # synthetic stand-in keeps its shape statistics: 24 rows, 250 Hz, T=625.
    "cfg1": SynthConfig("cfg1", (71, 64, 57), 24, 625, "classic", sampling_rate=250,
                        node_spacing_km=0.025, vp=3.6, vs=3.6 / 1.9, sta=0.01, lta=0.25,
                        seed=1, n_events=1),
    "cfg2": SynthConfig("cfg2", (101, 101, 51), 32, 120000, "classic", seed=2, n_events=40),
    "cfg3": SynthConfig("cfg3", (201, 201, 101), 64, 360000, "classic", seed=3, n_events=100),
    "cfg4": SynthConfig("cfg4", (151, 151, 151), 48, 30000, "centred", seed=4, n_events=1),
    "cfg5": SynthConfig("cfg5", (301, 301, 151), 128, 600000, "centred", seed=5, n_events=150),
    # the Iceland example geometry scanned at 50 Hz (tests/test_benchmarks.py serves the LUT at 50 Hz):
    # 25 m nodes are ~0.35 samples apart, so most neighbouring nodes share travel-time offsets
    "cfg1fine": SynthConfig("cfg1fine", (71, 64, 57), 24, 30000, "classic", sampling_rate=50,
                            node_spacing_km=0.025, vp=3.6, vs=3.6 / 1.9, sta=0.05, lta=0.5,
                            seed=1, n_events=10),
    # one GPU's x-slab of cfg5 when the grid is sharded over 8 GPUs (38 of 301 x-planes)
    "cfg5slab": SynthConfig("cfg5slab", (38, 301, 151), 128, 600000, "centred", seed=5, n_events=150),
}


@dataclasses.dataclass
class SynthProblem:
    cfg: SynthConfig
    traveltimes: np.ndarray      # int32 (nx, ny, nz, S)
    fsmp: int
    lsmp: int
    nsta: int
    nlta: int
    availability: dict           # "STATION_PHASE" -> 1, insertion-ordered like stalta.py:413
    row_channels: list           # per onset row: list of channel indices into ``waveforms``
    waveforms: np.ndarray | None  # float64 (C, P) or None when only geometry was requested
    stations_xyz: np.ndarray     # (n_stations, 3) km
    events: list                 # [(flat_node, origin_sample_in_output_frame)]

    @property
    def n_total(self) -> int:    # P = fsmp + T + lsmp
        return self.fsmp + self.cfg.n_samples + self.lsmp


def _layout_rows(n_rows: int):
    """Rows are phase-outer / station-inner (stalta.py:389,413).  Odd S -> P only."""
    if n_rows % 2 == 0:
        n_sta = n_rows // 2
        return n_sta, ["P"] * n_sta + ["S"] * n_sta
    return n_rows, ["P"] * n_rows


def make_geometry(cfg: SynthConfig):
    rng = np.random.default_rng(cfg.seed)
    n_sta, phases = _layout_rows(cfg.n_rows)
    nx, ny, nz = cfg.grid
    h = cfg.node_spacing_km
    xyz = np.column_stack([
        rng.uniform(0.0, (nx - 1) * h, n_sta),
        rng.uniform(0.0, (ny - 1) * h, n_sta),
        np.zeros(n_sta),
    ])
    return n_sta, phases, xyz


def make_traveltimes(cfg: SynthConfig, x_range=None) -> tuple[np.ndarray, np.ndarray, float]:
    """
    int32 sample-offset LUT ``(nx, ny, nz, S)`` = ``rint(tt * sampling_rate)`` (lut.py:538).
    ``x_range=(i0, i1)`` builds only an x-slab (contiguous flat-index range) -- what one rank
    of a node-sharded run needs.  Returns (lut, stations_xyz, max_traveltime_seconds_full_grid).
    """
    n_sta, phases, xyz = make_geometry(cfg)
    nx, ny, nz = cfg.grid
    h = cfg.node_spacing_km
    i0, i1 = (0, nx) if x_range is None else x_range
    gx = (np.arange(i0, i1) * h)[:, None, None]
    gy = (np.arange(ny) * h)[None, :, None]
    gz = (np.arange(nz) * h)[None, None, :]
    lut = np.empty((i1 - i0, ny, nz, cfg.n_rows), dtype=np.int32)
    for r in range(cfg.n_rows):
        sx, sy, sz = xyz[r % n_sta]
        v = cfg.vp if phases[r] == "P" else cfg.vs
        dist = np.sqrt((gx - sx) ** 2 + (gy - sy) ** 2 + (gz - sz) ** 2)
        lut[..., r] = np.rint(dist / v * cfg.sampling_rate).astype(np.int32)
    # max travel time over the FULL grid (corner distances bound it exactly for a box)
    corners = np.array([[a, b, c] for a in (0.0, (nx - 1) * h) for b in (0.0, (ny - 1) * h)
                        for c in (0.0, (nz - 1) * h)])
    dmax = max(np.linalg.norm(corners - s, axis=1).max() for s in xyz)
    vmin = cfg.vs if "S" in phases else cfg.vp
    return lut, xyz, float(dmax / vmin)


def _wavelet(freq: float, sr: int) -> np.ndarray:
    """Gaussian-derivative wavelet, unit peak (cf. examples/synthetic/simulate/core.py:51)."""
    half = int(round(1.5 / freq * sr))
    t = np.arange(-half, half + 1) / sr
    a = (np.pi * freq) ** 2
    w = -t * np.exp(-a * t * t)
    return w / np.abs(w).max()


def make_problem(cfg: SynthConfig, with_waveforms: bool = True, x_range=None) -> SynthProblem:
    """Geometry + LUT (+ waveforms).  Seeds: geometry ``cfg.seed``, waveforms ``100 + cfg.seed``."""
    lut, xyz, ttmax = make_traveltimes(cfg, x_range=x_range)
    sr = cfg.sampling_rate
    pre_pad, post_pad = stalta_pads(cfg.sta, cfg.lta, ttmax, cfg.n_samples / sr, sr)
    fsmp, lsmp = time2sample(pre_pad, sr), time2sample(post_pad, sr)
    nsta, nlta = time2sample(cfg.sta, sr) + 1, time2sample(cfg.lta, sr) + 1  # stalta.py:395-396
    n_sta, phases, _ = make_geometry(cfg)
    total = fsmp + cfg.n_samples + lsmp

    # channels: station k has Z (index 3k), N (3k+1), E (3k+2); P rows use Z, S rows use N+E
    availability, row_channels = {}, []
    for r in range(cfg.n_rows):
        k = r % n_sta
        availability[f"ST{k:03d}_{phases[r]}"] = 1
        row_channels.append([3 * k] if phases[r] == "P" else [3 * k + 1, 3 * k + 2])

    events, waveforms = [], None
    if with_waveforms:
        rng = np.random.default_rng(100 + cfg.seed)
        waveforms = rng.standard_normal((3 * n_sta, total))
        w = _wavelet(4.0, sr)
        half = len(w) // 2
        nx, ny, nz = cfg.grid
        h = cfg.node_spacing_km
        gap = max(cfg.n_samples // max(cfg.n_events, 1), 1)
        for e in range(cfg.n_events):
            node = int(rng.integers(0, cfg.n_nodes))
            ix, iy, iz = np.unravel_index(node, cfg.grid)
            origin = int(e * gap + rng.integers(0, max(gap // 2, 1)))  # output-frame sample
            amp = rng.uniform(5.0, 20.0)
            pos = np.array([ix * h, iy * h, iz * h])
            dist = np.linalg.norm(xyz - pos, axis=1)
            for k in range(n_sta):
                for ch, v in ((3 * k, cfg.vp), (3 * k + 1, cfg.vs), (3 * k + 2, cfg.vs)):
                    arr = fsmp + origin + int(np.rint(dist[k] / v * sr))
                    a, b = arr - half, arr + half + 1
                    if a < 0 or b > total:
                        continue
                    waveforms[ch, a:b] += amp * w
            events.append((node, origin))
    return SynthProblem(cfg, lut, fsmp, lsmp, nsta, nlta, availability, row_channels,
                        waveforms, xyz, events)


def onsets_from_waveforms(problem: SynthProblem, stalta_fn, min_onset_value: float = 0.4):
    """
    The ``STALTAOnset._onset`` recipe (stalta.py:516-548) with a caller-supplied STA/LTA
    ``stalta_fn(signal, nsta, nlta) -> onset`` (the CUDA op in the product, the oracle in tests):
    energy transform -> per-channel STA/LTA -> RMS over channels -> clip(min_onset_value).
    Returns float64 ``(S, P)`` in row order of ``problem.availability``.
    """
    wf = problem.waveforms
    cache = {}
    rows = []
    for chans in problem.row_channels:
        per_ch = []
        for c in chans:
            if c not in cache:
                cache[c] = stalta_fn(wf[c] ** 2, problem.nsta, problem.nlta)
            per_ch.append(cache[c])
        onset = np.sqrt(np.sum([o ** 2 for o in per_ch], axis=0) / len(per_ch))
        rows.append(np.clip(onset, min_onset_value, np.inf))
    return np.stack(rows, axis=0)


def random_problem(seed: int, grid, n_rows: int, n_samples: int, fsmp: int, lsmp: int,
                   tau_max: int | None = None, neg_frac: float = 0.0, floor_frac: float = 0.2):
    """
    Unstructured stress input: random int32 tau in [0, tau_max] (optionally some negative, to
    exercise the clamp of migratelib.c:55), gamma-distributed linear onsets with a fraction
    sitting exactly on a clip floor (exact ties).  Returns (onsets f64 (S,P), tau int32 (..., S)).
    """
    rng = np.random.default_rng(seed)
    tau_max = lsmp if tau_max is None else tau_max
    tau = rng.integers(0, tau_max + 1, size=tuple(grid) + (n_rows,), dtype=np.int32)
    if neg_frac > 0:
        mask = rng.random(tau.shape) < neg_frac
        tau[mask] = -rng.integers(1, 50, size=int(mask.sum()), dtype=np.int32)
    total = fsmp + n_samples + lsmp
    onsets = rng.gamma(2.0, 0.6, size=(n_rows, total))
    floor = rng.random(onsets.shape) < floor_frac
    onsets[floor] = 0.4
    onsets = np.clip(onsets, 0.4, np.inf)
    return onsets, tau
