# -*- coding: utf-8 -*-
"""
Copy the reference's own golden LOCATE run of the Icequake_Iceland example -- inputs and outputs --
so that the whole chain (raw traces -> decimation -> detrend / taper / band-pass -> centred STA/LTA
-> migrate -> marginal window -> location fits) is pinned against numbers the reference produced
with real obspy / scipy / pyproj, on the GPU box where /root/reference does not exist.

Run in the dev container (needs /root/reference).  examples/benchmarks/Icequake_Iceland holds, per
located event, the raw cut waveforms the reference read for it (<uid>.m: miniSEED, 36 traces at
500 Hz covering exactly [trigger - 2 mw - pre_pad, trigger + 2 mw + post_pad]; written by
io/cut_waveforms.py:45-92 from event.data.raw_waveforms) and its results (<uid>.event,
io/event.py:441-505).  tests/test_benchmarks.py:141-160 of the reference compares locate() against
the latter.  Written here (bit-identical copies unless noted):

  iceland_locate/<uid>.m, <uid>.event           three events
  iceland_locate/iceland_stations.txt           examples/Icequake_Iceland/inputs/iceland_stations.txt
  iceland_locate/TriggeredEvents.csv            the trigger file locate() started from
  iceland_locate/lut_facts.json (generated)     how iceland_lut.py built example.LUT (homogeneous,
                                                vp / vs, corners, spacing, projections), and the proof
                                                that tests/iceland_example.py regenerates that table:
                                                max |tt - tt_pickle| and the count of int32 offsets
                                                that differ at 250 Hz (0)

    python tests/golden/make_golden_iceland_events.py
"""
import json
import pathlib
import shutil
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

HERE = pathlib.Path(__file__).resolve().parent
OUT = HERE / "iceland_locate"
EX = pathlib.Path("/root/reference/examples")
BENCH = EX / "benchmarks" / "Icequake_Iceland"

if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    for f in sorted(BENCH.glob("*.m")) + sorted(BENCH.glob("*.event")):
        shutil.copyfile(f, OUT / f.name)
    shutil.copyfile(EX / "Icequake_Iceland" / "inputs" / "iceland_stations.txt", OUT / "iceland_stations.txt")
    shutil.copyfile(BENCH / "example_run_2014_180_TriggeredEvents.csv", OUT / "TriggeredEvents.csv")
    facts = {
        # examples/Icequake_Iceland/iceland_lut.py:22-53
        "method": "homogeneous", "vp": 3.630, "vs": 1.833,
        "ll_corner_lonlatdepth": [-17.24, 64.322, -1.4], "ur_corner_lonlatdepth": [-17.204, 64.336, 0.0],
        "node_spacing": [0.025, 0.025, 0.025],
        "grid_proj": "+proj=lcc +units=km +lon_0=-17.222 +lat_0=64.329 +lat_1=64.323 +lat_2=64.335 +datum=WGS84 "
                     "+ellps=WGS84 +no_defs",
        # examples/Icequake_Iceland/iceland_locate.py:47-70
        "locate": {"position": "centred", "sampling_rate": 250, "phases": ["P", "S"],
                   "bandpass_filters": {"P": [10, 124, 4], "S": [10, 124, 4]},
                   "sta_lta_windows": {"P": [0.01, 0.25], "S": [0.05, 0.5]}, "marginal_window": 0.06},
    }
    (OUT / "lut_facts.json").write_text(json.dumps(facts, indent=1))
    import iceland_example as ice
    from quakemigrate_b200.lut import read_reference_lut
    mine = ice.build_lut()
    ref = read_reference_lut(BENCH / "example.LUT")
    assert list(mine.node_count) == list(ref.node_count)
    worst, differ = 0.0, 0
    for st in ref.stations:
        for ph in ("P", "S"):
            worst = max(worst, float(np.abs(mine[st][ph] - ref[st][ph]).max()))
            differ += int((np.rint(mine[st][ph] * 250) != np.rint(ref[st][ph] * 250)).sum())
    facts["regenerated_vs_pickle"] = {"max_abs_seconds": worst, "int32_offsets_differing_at_250Hz": differ,
                                      "ll_corner_grid": [float(v) for v in ref.ll_corner]}
    assert differ == 0 and worst < 1e-11
    (OUT / "lut_facts.json").write_text(json.dumps(facts, indent=1))
    print("wrote", OUT, facts["regenerated_vs_pickle"])
