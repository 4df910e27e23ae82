# -*- coding: utf-8 -*-
"""
Parity at the sizes that stress the error band: the candidate band of the
16-bit scan is delta = S * q_step / 2 and the number of near-tied nodes grows with the grid, so
BASELINE configs 3, 4 and 5 get the same treatment config 2 has in test_gpu_fullsize.py --
oracle spot windows on the FULL grid (arg-max bit-exact, max_coa 1e-9, max_norm_coa 1e-5; the
oracle evaluates 200 samples of the full grid, seconds on the host cores), planted events,
sharding invariance -- at grid sizes of the configs and time spans short enough for a test:

* cfg3: 201x201x101 nodes (4.08 M), 64 rows, classic, T = 20000
* cfg5: one GPU's slab of the 8-way split, 38x301x151 nodes (1.73 M), 128 rows, centred, T = 8000
* cfg4: 151^3 nodes (3.44 M), 48 rows, centred, T = 6000; the marginal-window slab against
  migrate() of the reference restatement on that window (scan.py:719-721, io/event.py:433-435)
"""

import numpy as np
import pytest

from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn
from quakemigrate_b200.sharding import merge_rank_results, shard_traveltimes

pytestmark = pytest.mark.gpu

CASES = {
    "cfg3": ("cfg3", None, 20000),
    "cfg5slab": ("cfg5slab", None, 8000),
    "cfg4": ("cfg4", None, 6000),
}


def _setup(name):
    from quakemigrate_b200.core import Engine
    base, grid, T = CASES[name]
    import dataclasses
    # events 500 samples apart (the configs plant theirs for the full-length spans)
    cfg = dataclasses.replace(syn.CONFIGS[base].scaled(grid=grid, n_samples=T, name=name), n_events=max(1, T // 500))
    prob = syn.make_problem(cfg)
    eng = Engine(0)
    dev, onsets = eng.onsets_from_waveforms(prob.waveforms, prob.row_channels, prob.nsta, prob.nlta,
                                            position=cfg.position, return_onsets=True)
    lut = eng.upload_lut(prob.traveltimes)
    ms = eng.detect_device(lut, dev, prob.fsmp, prob.lsmp, cfg.n_rows)
    info = eng.last_scan_info
    res = eng.detect_fetch(cfg.n_samples)
    return cfg, prob, eng, lut, dev, onsets, res, ms, info


@pytest.fixture(scope="module", params=list(CASES))
def run(request):
    out = _setup(request.param)
    yield out
    out[2].close()


def test_scan_used_verified_16bit_rows(run):
    cfg, prob, eng, lut, dev, onsets, res, ms, info = run
    assert info["operand_bytes"] == 2 and info["fell_back_fp32"] == 0
    assert info["check_windows"] >= 1 and info["check_max_rel_err"] <= 2e-6
    assert info["host_syncs_in_timed_region"] == 0
    # the slow path stays an exception at these sizes (band ~ S * q_step)
    assert info["slow_path_samples"] <= cfg.n_samples // 100


def test_oracle_spot_windows(run):
    cfg, prob, eng, lut, dev, onsets, (mx, nm, ix), ms, info = run
    rng = np.random.default_rng(11)
    n_win = 160
    starts = [0, cfg.n_samples - n_win, int(rng.integers(0, cfg.n_samples - n_win))]
    if prob.events:
        starts.append(int(np.clip(prob.events[0][1] - n_win // 2, 0, cfg.n_samples - n_win)))
    for a in starts:
        sub = np.ascontiguousarray(onsets[:, a:a + prob.fsmp + n_win + prob.lsmp])
        w_mx, w_nm, w_ix = O.detect_fused(sub, prob.traveltimes, prob.fsmp, prob.lsmp, cfg.n_rows,
                                          threads=O.default_threads())
        assert np.array_equal(ix[a:a + n_win], w_ix), f"{cfg.name}: window at {a}"
        np.testing.assert_allclose(mx[a:a + n_win], w_mx, rtol=1e-9)
        np.testing.assert_allclose(nm[a:a + n_win], w_nm, rtol=1e-5)


def test_planted_events_recovered(run):
    cfg, prob, eng, lut, dev, onsets, (mx, nm, ix), ms, info = run
    usable = [(n, o) for n, o in prob.events if 50 <= o < cfg.n_samples - 60]
    hits = 0
    for node, origin in usable:
        seg = slice(origin - 40, origin + 60)
        t = origin - 40 + int(np.argmax(mx[seg]))
        got = np.array(np.unravel_index(ix[t], cfg.grid))
        want = np.array(np.unravel_index(node, cfg.grid))
        if np.abs(got - want).max() <= 2 and mx[t] > 1.5 * np.median(mx):
            hits += 1
    assert hits >= 0.8 * len(usable)


def test_sharding_invariance(run):
    """Three x-slabs with their node offsets, folded on the host (what the ranks' exchange must
    reproduce), equal the whole-grid run: indices exactly."""
    cfg, prob, eng, lut, dev, onsets, (mx, nm, ix), ms, info = run
    parts = []
    for r in range(3):
        slab, off, n_glob = shard_traveltimes(prob.traveltimes, 3, r)
        sl = eng.upload_lut(slab, node_offset=off, n_nodes_global=n_glob)
        eng.detect_device(sl, dev, prob.fsmp, prob.lsmp, cfg.n_rows)
        a, b, c = eng.detect_fetch(cfg.n_samples)
        n_local = int(np.prod(slab.shape[:-1]))
        parts.append((a, b * n_local / n_glob, c, n_local))
        sl.free()
    g_mx, g_nm, g_ix = merge_rank_results(parts)
    assert np.array_equal(g_ix, ix)
    np.testing.assert_allclose(g_mx, mx, rtol=1e-12)
    np.testing.assert_allclose(g_nm, nm, rtol=2e-6)


def test_marginal_window_against_reference_migrate(run):
    """locate (cfg4's path, run for every size): the window around the strongest sample, as a slab
    and as its time-sum, against migrate() on exactly that window (float32-row kernel: 1e-5)."""
    cfg, prob, eng, lut, dev, onsets, (mx, nm, ix), ms, info = run
    imax = int(np.argmax(mx))
    w0 = int(np.clip(imax - 12, 0, cfg.n_samples - 25))
    w1 = w0 + 25
    coa, slab = eng.locate_window(lut, dev, prob.fsmp, prob.lsmp, cfg.n_rows, w0, w1, return_map4d=True)
    sub = np.ascontiguousarray(onsets[:, w0:w0 + prob.fsmp + (w1 - w0) + prob.lsmp])
    m4 = O.migrate(sub, prob.traveltimes, prob.fsmp, prob.lsmp, cfg.n_rows, threads=O.default_threads())
    np.testing.assert_allclose(slab, m4, rtol=1e-5)
    np.testing.assert_allclose(coa, m4.sum(axis=-1), rtol=1e-5)
    flat = slab.reshape(-1, w1 - w0)
    assert np.array_equal(np.argmax(flat, axis=0), ix[w0:w1])
    # the device-resident form gives the same map
    dmap = eng.locate_window_device(lut, dev, prob.fsmp, prob.lsmp, cfg.n_rows, w0, w1)
    np.testing.assert_array_equal(dmap.fetch(), coa)
    dmap.free()
