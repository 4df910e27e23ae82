# -*- coding: utf-8 -*-
"""
Node-sharded detect over NCCL (one process per GPU): every rank uploads its x-slab of the
travel-time table, all ranks must end with the global, oracle-identical result.
Needs >= 2 GPUs (skipped otherwise); run with ``gpurun --gpus 2``.
"""

import os
import pathlib
import subprocess
import sys
import time

import pytest

from quakemigrate_b200 import _lib

ROOT = pathlib.Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

_WORKER = r"""
import os, sys, time, pathlib
sys.path.insert(0, {root!r})
import numpy as np
from oracle import qm_oracle as O
from quakemigrate_b200 import synthetic as syn
from quakemigrate_b200.core import Engine
from quakemigrate_b200.sharding import shard_traveltimes

rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), pathlib.Path(sys.argv[3])
eng = Engine(rank)
if rank == 0:
    tmp = idfile.with_suffix(".tmp")
    tmp.write_bytes(Engine.new_comm_id())
    tmp.rename(idfile)
else:
    t0 = time.time()
    while not idfile.exists():
        assert time.time() - t0 < 120, "timed out waiting for the NCCL id"
        time.sleep(0.05)
eng.init_comm(idfile.read_bytes(), world, rank)

for case, (grid, S, T, fsmp, lsmp, floor) in enumerate([((9, 6, 5), 6, 300, 5, 30, 0.2),
                                                        ((8, 8, 8), 4, 260, 3, 25, 0.9),   # ties -> slow path
                                                        ((13, 7, 6), 17, 700, 9, 40, 0.1)]):
    onsets, tau = syn.random_problem(100 + case, grid, S, T, fsmp, lsmp, neg_frac=0.03, floor_frac=floor)
    slab, off, n_glob = shard_traveltimes(tau, world, rank)
    lut = eng.upload_lut(slab, node_offset=off, n_nodes_global=n_glob)
    got = eng.detect(lut, onsets, fsmp, lsmp, S)
    want = O.find_max_coa(O.migrate(onsets, tau, fsmp, lsmp, S, threads=2), threads=2)
    assert np.array_equal(got[2], want[2]), f"case {{case}}: argmax differs on rank {{rank}}"
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5)
    lut.free()
eng.close()
print("rank", rank, "ok")
"""


@pytest.mark.skipif(_lib.LIB_PATH.exists() and _lib.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_detect_two_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    idfile = tmp_path / "nccl.id"
    env = dict(os.environ, NCCL_DEBUG="WARN")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", str(idfile)], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    t0 = time.time()
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=max(10, 400 - (time.time() - t0)))[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append("TIMEOUT\n" + p.communicate()[0])
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
