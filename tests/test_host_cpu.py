# -*- coding: utf-8 -*-
"""
CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, host logic of the
plugin mirrors, pad arithmetic, sharding (incl. a world-size-2 gloo run).
"""

import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from quakemigrate_b200 import _lib, build

    lib = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib)], check=True, capture_output=True,
                         text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    declared = ge.declared_symbols()
    assert len(declared) >= 25
    assert not [s for s in declared if s not in exported]
    # ctypes can open it and every signature in the binding resolves (no compute call is made)
    cdll = _lib.load()
    assert set(_lib.SIGNATURES) == set(declared)
    assert cdll.qmb_version().startswith(b"qmb200")


def test_built_for_sm100a_with_tma_bulk_copies():
    """The product is sm_100a SASS with UBLKCP (cp.async.bulk) in the scan kernel."""
    from quakemigrate_b200 import build
    lib = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", str(lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_no_cpu_fallback():
    """Without a device the product path fails loudly instead of computing on the host."""
    from quakemigrate_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    from quakemigrate_b200.core import Engine, QmbError
    with pytest.raises(QmbError):
        Engine(0)


def test_pad_arithmetic_matches_reference_formulas():
    from quakemigrate_b200 import synthetic as syn
    from quakemigrate_b200.onsets import STALTAOnset
    # util.time2sample / trim2sample (util.py:152-172, 248-269)
    assert syn.time2sample(1.652, 250) == 413
    assert syn.trim2sample(1.6501, 250) == 1.652
    on = STALTAOnset(sampling_rate=250, sta_lta_windows={"P": [0.01, 0.25], "S": [0.05, 0.5]})
    assert on.pre_pad == pytest.approx(0.5 + 3 * 0.05)
    on.post_pad = 1.527       # lut.max_traveltime of the Iceland example (note 1)
    assert on.post_pad == 3.0
    pre, post = on.pad(2.5)   # iceland_detect.py timestep
    # SURVEY section 8 footnote: pre_pad 1.652 s -> fsmp 413; post_pad 4.0 s -> lsmp 1000
    assert (syn.time2sample(pre, 250), syn.time2sample(post, 250)) == (413, 1000)
    assert syn.stalta_pads(0.05, 0.5, 1.527, 2.5, 250) == (pre, post)


def test_lut_contract():
    from quakemigrate_b200.lut import LUT
    rng = np.random.default_rng(0)
    tts = {f"ST{k}": {"P": rng.random((3, 4, 5)), "S": rng.random((3, 4, 5)) * 2} for k in range(3)}
    lut = LUT(tts, (3, 4, 5), (0.5, 0.5, 0.25), (10.0, 20.0, -1.0))
    avail = {"ST0_P": 1, "ST1_P": 0, "ST2_P": 1, "ST0_S": 1, "ST1_S": 1, "ST2_S": 0}
    served = lut.serve_traveltimes(100, avail)
    assert served.dtype == np.int32 and served.shape == (3, 4, 5, 4)
    assert np.array_equal(served[..., 1], np.rint(tts["ST2"]["P"] * 100).astype(np.int32))
    assert np.array_equal(served[..., 3], np.rint(tts["ST1"]["S"] * 100).astype(np.int32))
    assert lut.serve_traveltimes(100).shape[-1] == 6          # phase-outer, station-inner
    assert lut.max_traveltime == max(v["S"].max() for v in tts.values())
    # index2coord(unravel=True): C-order, ll_corner + ijk * spacing (lut.py:165-166, 178)
    flat = np.array([0, 5, 59, 23])
    ijk = np.column_stack(np.unravel_index(flat, (3, 4, 5)))
    want = np.array([10.0, 20.0, -1.0]) + ijk * np.array([0.5, 0.5, 0.25])
    assert np.allclose(lut.index2coord(flat, unravel=True), want)


def test_marginal_window_indices():
    from quakemigrate_b200.scan import marginal_window_indices
    coa = np.ones(401)
    coa[150] = 3.0
    coa[300] = 3.0                      # first maximum wins (pandas idxmax, event.py:239)
    assert marginal_window_indices(coa, 0.5, 100) == (100, 200, 150)
    coa[:] = 1.0
    coa[10] = 2.0
    assert marginal_window_indices(coa, 0.5, 100) == (0, 60, 10)
    coa[:] = 1.0
    coa[395] = 2.0
    assert marginal_window_indices(coa, 0.5, 100) == (345, 400, 395)


def test_stalta_onset_plan_orders_rows_like_the_reference():
    from quakemigrate_b200.onsets import STALTAOnset, WaveformData
    n = 300
    tr = lambda k: np.sin(np.arange(n) * (0.1 + 0.01 * k))          # noqa: E731  (flat-lined traces are dropped)
    ch = {"A": {"Z": tr(0), "N": tr(1), "E": tr(2)},
          "B": {"Z": tr(3)},                                        # no horizontals -> no S row
          "C": {"N": tr(4), "E": tr(5)}}                            # no vertical -> no P row
    on = STALTAOnset(sampling_rate=100)
    avail, rows, chans = on._plan(WaveformData(ch, 0.0, 100))
    assert list(avail.items()) == [("A_P", 1), ("B_P", 1), ("C_P", 0), ("A_S", 1), ("B_S", 0), ("C_S", 1)]
    assert [(r[0], r[1]) for r in rows] == [("A", "P"), ("B", "P"), ("A", "S"), ("C", "S")]
    assert rows[0][3:] == (21, 101)                                  # nsta/nlta = round(s*sr)+1
    assert [len(r[2]) for r in rows] == [1, 1, 2, 2] and len(chans) == 6


def test_slab_bounds_partition():
    from quakemigrate_b200.sharding import slab_bounds
    for nx, world in [(101, 8), (8, 8), (201, 3), (7, 1)]:
        edges = [slab_bounds(nx, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == nx
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab_bounds(3, 4, 0)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn
from quakemigrate_b200.sharding import shard_traveltimes, merge_rank_results

dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
grid, S, T, fsmp, lsmp = (9, 5, 4), 6, 120, 5, 30
onsets, tau = syn.random_problem(21, grid, S, T, fsmp, lsmp, floor_frac=0.7)   # many exact ties
slab, off, n_glob = shard_traveltimes(tau, world, rank)
m4 = O.migrate(onsets, slab, fsmp, lsmp, S)
mx, nm, ix = O.find_max_coa(m4, 1)
mine = (mx, nm, ix + off, int(np.prod(slab.shape[:-1])))
gathered = [None] * world
dist.all_gather_object(gathered, mine)
got = merge_rank_results(gathered)
want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S), 1)
assert n_glob == int(np.prod(grid))
assert np.array_equal(got[2], want[2]), "argmax"
assert np.array_equal(got[0], want[0]), "max"
np.testing.assert_allclose(got[1], want[1], rtol=1e-13)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_merge_world2_gloo(tmp_path):
    """Node-slab sharding + candidate fold reproduce the single-process result (2 ranks, gloo)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_lut_decimate_matches_reference_slicing():
    """LUT.decimate (lut/lut.py:102-140): node count, spacing and the centred start offset."""
    from quakemigrate_b200.lut import LUT

    rng = np.random.default_rng(5)
    shape = (11, 8, 7)
    tts = {f"S{k}": {"P": rng.random(shape), "S": rng.random(shape)} for k in range(3)}
    lut = LUT(tts, shape, (0.5, 0.5, 0.25), (1.0, 2.0, -3.0))
    dec = lut.decimate([2, 3, 2])
    assert list(dec.node_count) == [6, 3, 4] and list(lut.node_count) == [11, 8, 7]
    np.testing.assert_allclose(dec.node_spacing, [1.0, 1.5, 0.5])
    np.testing.assert_allclose(dec.ll_corner, lut.ll_corner)  # the reference leaves it untouched
    # c1 = (n - df * (new - 1) - 1) // 2 = (0, 0, 0) for x, (8 - 6 - 1) // 2 = 0 for y, 0 for z
    np.testing.assert_array_equal(dec["S1"]["S"], tts["S1"]["S"][0::2, 0::3, 0::2])
    dec2 = lut.decimate([4, 4, 4])
    # x: new = 3, c1 = (11 - 8 - 1) // 2 = 1;  y: new = 2, c1 = (8 - 4 - 1) // 2 = 1;  z: new = 2, c1 = 1
    np.testing.assert_array_equal(dec2["S0"]["P"], tts["S0"]["P"][1::4, 1::4, 1::4])
    assert dec2.serve_traveltimes(50).shape == (3, 2, 2, 6)
    lut.decimate([2, 2, 2], inplace=True)
    assert list(lut.node_count) == [6, 4, 4] and lut["S2"]["P"].shape == (6, 4, 4)


def test_balanced_ranges_partition_the_node_list():
    """sharding.balanced_range: contiguous, disjoint, covering, whole z-columns, sizes within one column."""
    from quakemigrate_b200.sharding import balanced_range, shard_traveltimes_balanced
    for (nx, ny, nz), world in (((201, 201, 101), 8), ((301, 301, 151), 8), ((7, 5, 3), 4), ((3, 3, 2), 9)):
        end, sizes = 0, []
        for r in range(world):
            first, count = balanced_range(nx, ny, nz, world, r)
            assert first == end and first % nz == 0 and count % nz == 0 and count > 0
            end = first + count
            sizes.append(count)
        assert end == nx * ny * nz and max(sizes) - min(sizes) <= nz
    tt = np.arange(4 * 3 * 2 * 5, dtype=np.int32).reshape(4, 3, 2, 5)
    got = [shard_traveltimes_balanced(tt, 5, r) for r in range(5)]
    assert np.array_equal(np.concatenate([g[0] for g in got]), tt.reshape(-1, 5))
    assert [g[1] for g in got] == [0, 6, 12, 16, 20] and all(g[2] == 24 for g in got)
    with pytest.raises(ValueError):
        balanced_range(2, 2, 3, 5, 0)
