# -*- coding: utf-8 -*-
"""bench.py contract pieces that can be checked without a GPU: the reference arm's JSON line."""

import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["unit"] == "node-samples/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert "workload" in j["config"] and j["value"] > 0


def test_reference_arm_other_ranks_exit_silently():
    env = {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    import os
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=str(ROOT), env=dict(os.environ, **env))
    assert out.returncode == 0 and out.stdout.strip() == ""
