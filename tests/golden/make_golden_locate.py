# -*- coding: utf-8 -*-
"""
Golden vectors for the location-fit stage that consumes the marginalised coalescence map
(SURVEY.md section 8f rank 1), produced by THE REFERENCE'S OWN CODE.

``quakemigrate`` cannot be imported here (obspy / pyproj / matplotlib are absent), so the five
methods ``QuakeScan._gaufilt3d``, ``_gaufit3d``, ``_covfit3d``, ``_mask3d``, ``_splineloc``
(quakemigrate/signal/scan.py:736-1070) and ``util.gaussian_3d`` (quakemigrate/util.py:76-116) are
lifted out of the reference's source files with ``ast`` -- unmodified, only the ``@util.timeit``
decorators dropped -- and executed against a stub ``self.lut`` whose projection is the identity.

    python tests/golden/make_golden_locate.py      # needs /root/reference
"""
import ast
import logging
import pathlib
import types

import numpy as np
from scipy.interpolate import Rbf
from scipy.signal import fftconvolve

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path("/root/reference/quakemigrate")


def lift(path, names, in_class=None):
    tree = ast.parse(path.read_text())
    body = tree.body
    if in_class:
        body = [n for n in body if isinstance(n, ast.ClassDef) and n.name == in_class][0].body
    funcs = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    for f in funcs:
        f.decorator_list = []
    assert {f.name for f in funcs} == set(names), [f.name for f in funcs]
    return funcs


def reference_namespace():
    util = types.SimpleNamespace()
    ns_util = {"np": np}
    mod = ast.Module(body=lift(REF / "util.py", {"gaussian_3d"}), type_ignores=[])
    exec(compile(ast.fix_missing_locations(mod), "util.py", "exec"), ns_util)
    util.gaussian_3d = ns_util["gaussian_3d"]
    funcs = lift(REF / "signal" / "scan.py", {"_gaufilt3d", "_gaufit3d", "_covfit3d", "_mask3d", "_splineloc"},
                 in_class="QuakeScan")
    cls = ast.ClassDef(name="Ref", bases=[], keywords=[], body=funcs, decorator_list=[])
    if "type_params" in ast.ClassDef._fields:
        cls.type_params = []
    mod = ast.Module(body=[cls], type_ignores=[])
    ns = {"np": np, "fftconvolve": fftconvolve, "Rbf": Rbf, "logging": logging, "util": util}
    exec(compile(ast.fix_missing_locations(mod), "scan.py", "exec"), ns)
    return ns["Ref"]


class StubLUT:
    """Identity projection: input coordinates == grid coordinates."""

    def __init__(self, node_count, node_spacing, ll_corner):
        self.node_count = np.array(node_count)
        self.node_spacing = np.array(node_spacing, dtype=float)
        self.ll_corner = np.array(ll_corner, dtype=float)

    def index2coord(self, value):
        return np.vstack(self.ll_corner + np.array(value) * self.node_spacing)

    def coord2grid(self, value, inverse=False):
        return np.atleast_2d(np.array(value))


def synthetic_map(shape, seed, centre, widths):
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    r2 = sum(((a - c) / w) ** 2 for a, c, w in zip(g, centre, widths))
    m = 1.0 + 6.0 * np.exp(-0.5 * r2) + 0.15 * rng.random(shape)
    m += 1.5 * np.exp(-0.5 * (((g[0] - 3) / 1.5) ** 2 + ((g[1] - 4) / 2.0) ** 2 + ((g[2] - 3) / 1.2) ** 2))
    return m / np.nanmax(m)          # scan.py:721


def main():
    Ref = reference_namespace()
    out = {}
    for tag, shape, centre, widths in [("a", (21, 18, 17), (11.3, 8.6, 9.2), (2.5, 1.8, 2.2)),
                                       ("b", (16, 24, 11), (2.2, 20.1, 5.4), (1.4, 2.6, 1.9))]:  # peak near an edge
        lut = StubLUT(shape, (0.5, 0.25, 0.4), (10.0, -3.0, 1.0))
        ref = Ref()
        ref.lut = lut
        coa_map = synthetic_map(shape, 7, centre, widths)
        smoothed = ref._gaufilt3d(np.copy(coa_map))
        gau_loc, gau_unc = ref._gaufit3d(np.copy(smoothed))
        cov_loc, cov_unc = ref._covfit3d(np.copy(coa_map))
        spl_loc = ref._splineloc(np.copy(coa_map))
        out.update({f"{tag}_coa_map": coa_map, f"{tag}_smoothed": smoothed,
                    f"{tag}_gau_loc": np.array(gau_loc), f"{tag}_gau_unc": np.array(gau_unc),
                    f"{tag}_cov_loc": np.array(cov_loc), f"{tag}_cov_unc": np.array(cov_unc),
                    f"{tag}_spl_loc": np.array(spl_loc),
                    f"{tag}_spacing": lut.node_spacing, f"{tag}_ll": lut.ll_corner})
        print(tag, "gau", gau_loc, gau_unc, "cov", cov_loc, cov_unc, "spline", spl_loc)
    np.savez_compressed(HERE / "golden_locate.npz", **out)
    print((HERE / "golden_locate.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
