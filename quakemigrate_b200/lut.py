# -*- coding: utf-8 -*-
"""
The part of the LUT plugin contract the coalescence hot path consumes, mirrored from
``quakemigrate/lut/lut.py``:

* ``LUT.serve_traveltimes(sampling_rate, availability)`` (lut.py:502-538): iterate the
  availability dict in order, stack the float64 ``(nx, ny, nz)`` travel-time grids of the
  available ``station_phase`` pairs along the last axis, ``rint(tt * sr).astype(int32)``;
* ``LUT.max_traveltime`` (lut.py:842-845), ``LUT.unit_conversion_factor``;
* ``Grid3D.index2grid`` / ``index2coord(value, unravel=True)`` (lut.py:142-243): C-order
  unravel, ``ll_corner + ijk * node_spacing``.  The pyproj inverse projection of the reference
  is geodesy outside the hot path: pass ``grid2coord`` (callable on an (n, 3) array) to apply
  one, otherwise coordinates stay in the grid projection.

LUT *construction* (create_lut.py: eikonal solvers, NonLinLoc) is out of scope; tables come from
the reference's pickles (:func:`read_reference_lut`, needs no pyproj) or from arrays.
"""

from __future__ import annotations

import pickle

import numpy as np


class Grid3D:
    def __init__(self, node_count, node_spacing, ll_corner, grid2coord=None):
        self.node_count = np.asarray(node_count, dtype=int)
        self.node_spacing = np.asarray(node_spacing, dtype=float)
        self.ll_corner = np.asarray(ll_corner, dtype=float)
        self.grid2coord = grid2coord
        self.unit_conversion_factor = 1.0

    def index2grid(self, value, inverse=False, unravel=False):
        """
        Grid index <-> grid coordinate (contract of lut.py:142-180): ``value`` is one ``[i, j, k]``,
        an ``(n, 3)`` array of them, or -- with ``unravel`` -- flat C-order node indices (x slowest,
        z fastest, lut.py:165-166).  Always returns an ``(n, 3)`` array; the forward map multiplies,
        then adds (the order the golden X/Y/Z integers of the reference's .scanmseed depend on).
        """
        idx = np.asarray(value)
        if unravel:
            idx = np.stack(np.unravel_index(idx.reshape(-1), tuple(int(n) for n in self.node_count)), axis=-1)
        idx = np.atleast_2d(idx)
        if inverse:
            return np.rint((idx - self.ll_corner) / self.node_spacing).astype(int)
        return self.ll_corner + idx * self.node_spacing

    def index2coord(self, value, inverse=False, unravel=False):
        """lut.py:211-243 with the projection step delegated to ``grid2coord``."""
        if inverse:
            raise NotImplementedError("coord -> index needs the forward projection (out of scope)")
        out = self.index2grid(value, unravel=unravel)
        return self.grid2coord(out) if self.grid2coord is not None else out


class LUT(Grid3D):
    """Travel-time look-up tables: ``lut[station][phase] -> float64 (nx, ny, nz)`` seconds."""

    def __init__(self, traveltimes: dict, node_count, node_spacing, ll_corner, phases=None,
                 grid2coord=None):
        super().__init__(node_count, node_spacing, ll_corner, grid2coord)
        self.traveltimes = traveltimes
        self.phases = phases or sorted({p for v in traveltimes.values() for p in v})
        self.stations = list(traveltimes.keys())

    def __getitem__(self, key):
        return self.traveltimes[key]

    def _grid(self, station, phase):
        """One travel-time grid; NonLinLoc imports store it under ``TIME_<phase>`` (lut.py:531-534)."""
        grids = self.traveltimes[station]
        return grids[phase] if phase in grids else grids[f"TIME_{phase}"]  # KeyError: phase missing

    def _stack(self, pairs):
        pairs = list(pairs)
        out = np.empty(tuple(int(n) for n in self.node_count) + (len(pairs),), dtype=np.double)
        for row, (station, phase) in enumerate(pairs):
            out[..., row] = self._grid(station, phase)
        return out

    def serve_traveltimes(self, sampling_rate, availability=None):
        """
        Contract of lut.py:502-538: the int32 sample-offset stack ``(nx, ny, nz, n_rows)``, one row
        per AVAILABLE ``"<station>_<phase>"`` key in the order of the availability dict (every
        station/phase, phase-outer, without one), ``rint(seconds * sampling_rate)``.
        """
        if availability is None:
            pairs = [(station, phase) for phase in self.phases for station in self.stations]
        else:
            pairs = [tuple(key.rsplit("_", 1)) for key, available in availability.items() if available == 1]
        return np.rint(self._stack(pairs) * sampling_rate).astype(np.int32)

    def _serve_traveltimes(self, phases, stations=None):
        """Float64 stack for the given phases, phase-outer / station-inner (contract of lut.py:576-603)."""
        stations = self.stations if stations is None else stations
        return self._stack((station, phase) for phase in phases for station in stations)

    def decimate(self, df, inplace=False):
        """
        lut.py:102-140: keep every ``df``-th node per axis, starting ``c1`` nodes in so that the
        kept nodes are centred.  Like the reference, ``node_spacing`` is scaled and ``ll_corner``
        is left as it is (the reference does not shift it by ``c1 * node_spacing``).
        """
        import copy

        df = np.array(df, dtype=int)
        new_node_count = 1 + (self.node_count - 1) // df
        c1 = (self.node_count - df * (new_node_count - 1) - 1) // 2
        grid = self if inplace else copy.copy(self)
        if not inplace:
            grid.traveltimes = {st: dict(m) for st, m in self.traveltimes.items()}
        grid.node_count = new_node_count
        grid.node_spacing = self.node_spacing * df
        for station, map_ in grid.traveltimes.items():
            for phase, ttimes in map_.items():
                map_[phase] = ttimes[c1[0]::df[0], c1[1]::df[1], c1[2]::df[2]]
        if not inplace:
            return grid

    @property
    def max_traveltime(self):
        """lut.py:842-845."""
        return np.max(self._serve_traveltimes(self.phases))

    # -- device-resident serving (the reference re-stacks and re-rounds on every call) --------
    def grid_index(self):
        """``{(station, phase): i}`` in the phase-outer / station-inner order of lut.py:576-603."""
        index = {}
        for phase in self.phases:
            for station in self.stations:
                tts = self[station]
                if phase in tts or f"TIME_{phase}" in tts:
                    index[(station, phase)] = len(index)
        return index

    def to_device(self, engine):
        """Upload every float64 grid once; returns ``(DeviceGrids, grid_index)``."""
        index = self.grid_index()
        stack = np.stack([self[st][ph] if ph in self[st] else self[st][f"TIME_{ph}"]
                          for (st, ph) in index], axis=0)
        return engine.upload_traveltime_grids(stack), index

    def decimated_grid(self, df) -> "Grid3D":
        """The grid of ``self.decimate(df)`` without touching the tables (lut.py:123-128): for use
        with ``DeviceGrids.decimate(df)``, which does the gather on the device."""
        df = np.broadcast_to(np.asarray(df, dtype=int), (3,))
        g = Grid3D(1 + (self.node_count - 1) // df, self.node_spacing * df, self.ll_corner, self.grid2coord)
        g.unit_conversion_factor = self.unit_conversion_factor
        return g

    @staticmethod
    def rows_for(availability, index):
        """Grid index per served row for an availability dict (lut.py:527-537 ordering)."""
        rows = []
        for key, available in availability.items():
            station, phase = key.split("_")
            if available == 1:
                rows.append(index[(station, phase)])   # KeyError -> missing phase, like the reference
        return rows


# ------------------------------------------------------------------------------------------------
# the one map projection the reference's examples use, so that ``index2coord`` can finish its job
# (lut.py:236: pyproj transform grid -> lon/lat) without pyproj
# ------------------------------------------------------------------------------------------------
class LambertConformalConic:
    """
    Ellipsoidal Lambert conformal conic with two standard parallels (``+proj=lcc``; Snyder,
    "Map projections -- a working manual", USGS PP 1395, eqs. 15-1..15-11, 14-15, 7-9) -- the
    published algorithm PROJ implements.  ``forward`` / ``inverse`` act on (n, 3) arrays
    ``[lon, lat, z]`` <-> ``[x, y, z]`` (z passes through), i.e. they play ``coord2grid`` /
    ``grid2coord`` (lut.py:182-243) for a ``longlat`` coordinate projection on the same datum.
    Pinned by the X / Y channels of the reference's golden .scanmseed (tests/test_scanmseed.py).
    """

    ELLIPSOIDS = {"WGS84": (6378137.0, 1.0 / 298.257223563), "GRS80": (6378137.0, 1.0 / 298.257222101)}
    UNITS = {"m": 1.0, "km": 1000.0}

    def __init__(self, lon_0, lat_0, lat_1, lat_2=None, ellps="WGS84", units="m", x_0=0.0, y_0=0.0):
        self.a, f = self.ELLIPSOIDS[ellps]
        self.e = np.sqrt(2.0 * f - f * f)
        self.unit = self.UNITS[units]
        self.unit_conversion_factor = self.unit  # pyproj axis_info[0].unit_conversion_factor
        self.lon_0, self.x_0, self.y_0 = np.radians(lon_0), x_0, y_0
        phi0, phi1 = np.radians(lat_0), np.radians(lat_1)
        phi2 = phi1 if lat_2 is None else np.radians(lat_2)
        if abs(phi1 - phi2) < 1e-12:
            self.n = np.sin(phi1)
        else:
            self.n = (np.log(self._m(phi1)) - np.log(self._m(phi2))) / (
                np.log(self._t(phi1)) - np.log(self._t(phi2)))
        self.F = self._m(phi1) / (self.n * self._t(phi1) ** self.n)
        self.rho0 = self.a * self.F * self._t(phi0) ** self.n

    def _m(self, phi):
        return np.cos(phi) / np.sqrt(1.0 - (self.e * np.sin(phi)) ** 2)

    def _t(self, phi):
        es = self.e * np.sin(phi)
        return np.tan(np.pi / 4 - phi / 2) / ((1.0 - es) / (1.0 + es)) ** (self.e / 2)

    def forward(self, coord):
        coord = np.atleast_2d(np.asarray(coord, dtype=float))
        lam, phi = np.radians(coord[:, 0]), np.radians(coord[:, 1])
        rho = self.a * self.F * self._t(phi) ** self.n
        theta = self.n * (lam - self.lon_0)
        out = coord.copy()
        out[:, 0] = (rho * np.sin(theta) + self.x_0) / self.unit
        out[:, 1] = (self.rho0 - rho * np.cos(theta) + self.y_0) / self.unit
        return out

    def inverse(self, grid):
        grid = np.atleast_2d(np.asarray(grid, dtype=float))
        x = grid[:, 0] * self.unit - self.x_0
        y = self.rho0 - (grid[:, 1] * self.unit - self.y_0)
        sgn = np.sign(self.n)
        rho = sgn * np.hypot(x, y)
        theta = np.arctan2(sgn * x, sgn * y)
        t = (rho / (self.a * self.F)) ** (1.0 / self.n)
        phi = np.pi / 2 - 2 * np.arctan(t)
        for _ in range(12):  # Snyder 7-9; converges to 1e-15 in ~5 rounds
            es = self.e * np.sin(phi)
            phi = np.pi / 2 - 2 * np.arctan(t * ((1.0 - es) / (1.0 + es)) ** (self.e / 2))
        out = grid.copy()
        out[:, 0] = np.degrees(theta / self.n + self.lon_0)
        out[:, 1] = np.degrees(phi)
        return out

    @classmethod
    def from_proj_string(cls, srs: str):
        """Parse a PROJ string (``+proj=lcc +units=km +lon_0=...``); None if it is not this projection."""
        kv = dict(tok[1:].split("=", 1) if "=" in tok else (tok[1:], "")
                  for tok in srs.split() if tok.startswith("+"))
        if kv.get("proj") != "lcc":
            return None
        ellps = kv.get("ellps", kv.get("datum", "WGS84"))
        if ellps not in cls.ELLIPSOIDS or kv.get("units", "m") not in cls.UNITS:
            return None
        return cls(float(kv.get("lon_0", 0.0)), float(kv.get("lat_0", 0.0)), float(kv["lat_1"]),
                   float(kv["lat_2"]) if "lat_2" in kv else None, ellps, kv.get("units", "m"),
                   float(kv.get("x_0", 0.0)), float(kv.get("y_0", 0.0)))


# ------------------------------------------------------------------------------------------------
# reading the reference's pickled LUTs without pyproj
# ------------------------------------------------------------------------------------------------
class _Opaque:
    """Stand-in for classes we cannot (and need not) import, e.g. pyproj.Proj."""

    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k

    def __setstate__(self, state):
        self.__dict__["state"] = state

    def __reduce_ex__(self, protocol):  # pragma: no cover
        return (_Opaque, ())


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if root in ("pyproj", "quakemigrate", "matplotlib", "obspy"):
            return _Opaque
        return super().find_class(module, name)


def read_reference_lut(path, grid2coord=None) -> LUT:
    """
    Load a LUT saved by the reference (``LUT.save``: pickle of ``LUT.__dict__``, lut.py:605-620;
    loader ``quakemigrate/io/core.py:24-51``) without importing pyproj/quakemigrate.
    """
    with open(path, "rb") as f:
        d = _StubUnpickler(f).load()
    if not isinstance(d, dict):
        d = getattr(d, "__dict__", {}).get("state", d)
    tt = d.get("traveltimes", d.get("maps"))
    node_count = d.get("_node_count", d.get("node_count"))
    spacing = d.get("node_spacing", d.get("_node_spacing", d.get("cell_size")))
    ll = d.get("ll_corner", d.get("_ll_corner"))
    if tt is None or node_count is None:
        raise ValueError(f"unrecognised LUT pickle (keys: {sorted(d)[:20]})")
    lut = LUT(tt, node_count, spacing if spacing is not None else np.ones(3),
              ll if ll is not None else np.zeros(3), phases=d.get("phases"), grid2coord=grid2coord)
    # grid projection: rebuilt from the PROJ string pyproj pickles (Proj.__reduce__ -> (srs,)) when it
    # is the Lambert conformal conic of the reference's examples; otherwise left to the caller
    gproj, cproj = d.get("grid_proj"), d.get("coord_proj")
    gargs, cargs = getattr(gproj, "args", ()), getattr(cproj, "args", ())
    if gargs and isinstance(gargs[0], str):
        lut.grid_proj_string = gargs[0]
        proj = LambertConformalConic.from_proj_string(gargs[0])
        if proj is not None:
            lut.unit_conversion_factor = proj.unit_conversion_factor
            if grid2coord is None and cargs and "+proj=longlat" in str(cargs[0]):
                lut.grid2coord, lut.coord2grid_fn = proj.inverse, proj.forward
    sd = d.get("station_data")
    if sd is not None and hasattr(sd, "__getitem__"):
        try:
            lut.stations = list(sd["Name"].values)
        except Exception:
            pass
    return lut
