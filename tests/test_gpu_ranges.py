# -*- coding: utf-8 -*-
"""
Slabs given as ranges of the C-order node list (qmb_lut_upload_range): a rank's share may begin and
end inside an x-plane, so that N / G nodes per rank are possible whatever nx is.  Folding the
per-range results (sharding.merge_rank_results: what the ranks' exchange must reproduce) gives the
whole-grid result: indices exactly, max_coa to rounding, the normalised maximum to 2e-6.
"""
import numpy as np
import pytest

from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn
from quakemigrate_b200.sharding import balanced_range, merge_rank_results, shard_traveltimes_balanced

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("grid,world", [((13, 11, 9), 3), ((21, 17, 12), 8), ((5, 7, 6), 5), ((2, 9, 8), 7)])
def test_balanced_ranges_reproduce_the_whole_grid(grid, world):
    from quakemigrate_b200.core import Engine
    S, T, fsmp, lsmp = 9, 700, 12, 60
    onsets, tau = syn.random_problem(17, grid, S, T, fsmp, lsmp, neg_frac=0.02, floor_frac=0.3)
    want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S, threads=4), threads=4)
    with Engine(0) as eng:
        dev = eng.upload_onsets(onsets)
        parts, covered = [], 0
        for r in range(world):
            rows, first, n_glob = shard_traveltimes_balanced(tau, world, r)
            lut = eng.upload_lut_range(rows, grid[1], grid[2], first, n_glob)
            info = lut.info()
            assert info["n_nodes_local"] == len(rows) and info["node_offset"] == first
            eng.detect_device(lut, dev, fsmp, lsmp, S)
            a, b, c = eng.detect_fetch(T)
            assert c.min() >= first and c.max() < first + len(rows)   # only nodes of this range can win
            parts.append((a, b * len(rows) / n_glob, c, len(rows)))
            covered += len(rows)
            # the marginal map of a range: one value per node of the range, in node order
            coa = eng.locate_window(lut, dev, fsmp, lsmp, S, 100, 140)
            m4 = O.migrate(np.ascontiguousarray(onsets[:, 100:100 + fsmp + 40 + lsmp]), tau, fsmp, lsmp, S, threads=4)
            np.testing.assert_allclose(coa.reshape(-1), m4.reshape(-1, 40)[first:first + len(rows)].sum(axis=-1),
                                       rtol=1e-5)
            lut.free()
        assert covered == int(np.prod(grid))
    g_mx, g_nm, g_ix = merge_rank_results(parts)
    assert np.array_equal(g_ix, want[2])
    np.testing.assert_allclose(g_mx, want[0], rtol=1e-9)
    np.testing.assert_allclose(g_nm, want[1], rtol=1e-5)


def test_range_arguments_are_checked():
    from quakemigrate_b200.core import Engine, QmbError
    tau = np.zeros((4 * 5 * 6, 3), dtype=np.int32)
    with Engine(0) as eng:
        with pytest.raises(QmbError):
            eng.upload_lut_range(tau[:7], 5, 6, 0, 120)        # not whole z-columns
        with pytest.raises(QmbError):
            eng.upload_lut_range(tau[:12], 5, 6, 3, 120)       # start inside a z-column
        with pytest.raises(QmbError):
            eng.upload_lut_range(tau[:60], 5, 6, 90, 120)      # past the end of the grid
        lut = eng.upload_lut_range(tau[:18], 5, 6, 12, 120)    # rows 2..4 of plane 0
        dev = eng.upload_onsets(np.ones((3, 200)))
        with pytest.raises(QmbError):
            eng.locate_window_device(lut, dev, 5, 10, 3, 0, 20)  # device maps need whole planes
        assert balanced_range(4, 5, 6, 2, 1) == (60, 60)
