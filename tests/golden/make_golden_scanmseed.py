# -*- coding: utf-8 -*-
"""
Copy the reference's own golden detect output and the facts of its LUT the output stage needs.

Run in the dev container (needs /root/reference):

  golden_2014_180.scanmseed   bit-identical copy of
                              examples/benchmarks/Icequake_Iceland/2014_180.scanmseed (the file
                              tests/test_benchmarks.py:80-105 compares detect() against): 5 channels
                              x 2500 samples, 11 STEIM-2 records written by obspy/libmseed
  golden_2014_180_StationAvailability.csv   bit-identical copy of the availability file of the same
                              run (tests/test_benchmarks.py:99-105 of the reference)
  golden_iceland_grid.json    node_count / node_spacing / ll_corner / PROJ strings of
                              examples/benchmarks/Icequake_Iceland/example.LUT

    python tests/golden/make_golden_scanmseed.py
"""
import json
import pathlib
import shutil
import sys

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from quakemigrate_b200.lut import _StubUnpickler  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent
BENCH = pathlib.Path("/root/reference/examples/benchmarks/Icequake_Iceland")

if __name__ == "__main__":
    shutil.copyfile(BENCH / "2014_180.scanmseed", HERE / "golden_2014_180.scanmseed")
    shutil.copyfile(BENCH / "2014_180_StationAvailability.csv", HERE / "golden_2014_180_StationAvailability.csv")
    with open(BENCH / "example.LUT", "rb") as f:
        d = _StubUnpickler(f).load()
    meta = {
        "node_count": [int(v) for v in d["_node_count"]],
        "node_spacing": [float(v) for v in d["_node_spacing"]],
        "ll_corner": [float(v) for v in d["ll_corner"]],
        "ur_corner": [float(v) for v in d["ur_corner"]],
        "grid_proj": d["grid_proj"].args[0],
        "coord_proj": d["coord_proj"].args[0],
    }
    (HERE / "golden_iceland_grid.json").write_text(json.dumps(meta, indent=1))
    print("wrote", HERE / "golden_2014_180.scanmseed", HERE / "golden_iceland_grid.json")
