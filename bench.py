# -*- coding: utf-8 -*-
"""
Benchmark of the coalescence hot path (fused migrate + stack + max-reduce) on B200.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own OpenMP C path on host cores
    python bench.py --workload cfg4           # locate path (mode "locate": span scan + marginal window + 3-D stage)
    python bench.py --workload cfg2 --weak    # round-1 workload: one cfg2 slab per GPU stacked along x

One "step" = one pass of the hot path over the workload: for every one of the N*T
(node, sample) pairs stack S onset samples, exponentiate, and reduce over nodes to the per-sample
(max, arg-max, normalised max) -- what QuakeScan._compute does per detect timestep
(quakemigrate/signal/scan.py:635-638).  Metric: node-samples / second (whole job).

Default workload: BASELINE.json configs[2] = north_star's target, 201x201x101 grid x 64 onset rows x
360000 samples @ 100 Hz, classic STA/LTA, GRID-SHARDED over the GPUs (strong scaling: N=1 holds the
whole grid, N=8 the survey's 8(e) split); every rank ends with the global result after one NCCL
all-gather of the per-sample candidates.  At every N, rank 0 runs the compiled reference C on a few
random 200-sample windows of the GLOBAL grid and the line carries the comparison
(`parity_vs_reference_sample`); those runs are also the `cpu_baseline`.

torch is used only for the torchrun rendezvous / barrier / max-over-ranks; the data path is
libqmb200.so (own CUDA kernels, own NCCL communicator) driven through ctypes.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from quakemigrate_b200 import synthetic as syn  # noqa: E402

CTRL_GROUP = None
METRIC = "grid-nodes x samples / sec (migrate+stack+max-reduce)"
UNIT = "node-samples/s"
DEFAULT_WORKLOAD = "cfg3"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(syn.CONFIGS))
    ap.add_argument("--mode", default=None, choices=["detect", "locate"],
                    help="detect: fused scan (default); locate: span scan + marginal window + 3-D map stage "
                         "(default for --workload cfg4)")
    ap.add_argument("--variant", type=int, default=-1, help="kernel tiling variant id; -1 = library default")
    ap.add_argument("--samples", type=int, default=0, help="override n_samples (debug)")
    ap.add_argument("--precision", default=None, choices=["auto", "q16", "fp32"],
                    help="operand format of the scan's gathers (library default: auto -> verified q16)")
    ap.add_argument("--qgroup", type=int, default=None, choices=[1, 2, 4, 8],
                    help="rows per carry-free group of the fixed-point scan (1: 16-bit rows ... 8: 13-bit rows); "
                         "default: the library's")
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the reference-C parity windows / CPU baseline (profiling runs)")
    ap.add_argument("--pre-only", action="store_true",
                    help="profiling aid: run only the pre-processing stage on the workload's full-length traces")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: one workload-sized slab per GPU stacked along x "
                         "(default: strong, the FIXED workload grid sharded over the ranks)")
    ap.add_argument("--strong", action="store_true", help="(default; kept for round-1 command lines)")
    ap.add_argument("--plane-slabs", action="store_true",
                    help="strong scaling with whole x-planes per rank (round-1 sharding) instead of equal node ranges")
    ap.add_argument("--e2e-steps", type=int, default=5,
                    help="cap of the end-to-end loop (host buffers in and out); the device-timed loop runs --steps")
    ap.add_argument("--parity-windows", type=int, default=3)
    ap.add_argument("--parity-samples", type=int, default=200)
    a = ap.parse_args()
    if a.mode is None:
        a.mode = "locate" if a.workload == "cfg4" else "detect"
    return a


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy burst)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def global_config(base: syn.SynthConfig, world: int, n_samples: int, weak: bool) -> syn.SynthConfig:
    nx, ny, nz = base.grid
    if not weak:
        return base.scaled(n_samples=n_samples or base.n_samples, name=base.name)
    return base.scaled(grid=(nx * world, ny, nz), n_samples=n_samples or base.n_samples,
                       name=f"{base.name}x{world}")


def git_head() -> str:
    try:
        return subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short=12", "HEAD"], capture_output=True,
                              text=True, timeout=5).stdout.strip() or "unknown"
    except Exception:
        return "unknown"


def ncu_traffic(workload: str, variant: int, operand_bits: int):
    """
    DRAM bytes per launch of k_scan from the committed `ncu --set full` capture of the SAME
    (workload, tiling variant, operand format): profiles/scan_ncu_summary.json holds one entry per
    key with the git revision it was captured at.  No entry -> null (never a number from another kernel).
    """
    prof = ROOT / "profiles" / "scan_ncu_summary.json"
    if not prof.exists():
        return None, None
    try:
        pj = json.loads(prof.read_text())
    except Exception:
        return None, None
    key = f"{workload}:v{variant}:" + ("fp32" if operand_bits >= 32 else f"q{operand_bits}")
    ent = (pj.get("entries") or {}).get(key)
    if not ent:
        return None, None
    return ent.get("dram_bytes_per_launch"), {"key": key, "captured_at_git": ent.get("git"), "tag": ent.get("tag")}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        if not self.rows:
            return None
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                          ("sw_power_cap", 7)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        pw = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own C path (oracle/_ref) or, where it was not built, the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_reference_window(traveltimes: np.ndarray, onsets_lin: np.ndarray, fsmp: int, lsmp: int, n_rows: int,
                         t0: int, t_sub: int):
    """
    migrate + find_max_coa of the reference (core/lib.py:53-170 semantics, incl. its fresh
    np.zeros volume) for output samples [t0, t0 + t_sub) of the FULL grid `traveltimes`.
    Returns ((max_coa, max_norm_coa, max_idx), seconds, kind, cores).
    """
    from oracle import qm_oracle as O  # bench.py touches oracle/ only in the CPU baseline / reference legs

    kind = "ref" if O.have("ref") else "port"
    cores = O.default_threads()
    sub = np.ascontiguousarray(onsets_lin[:, t0: t0 + fsmp + t_sub + lsmp])
    t_a = time.perf_counter()
    m4 = O.migrate(sub, traveltimes, fsmp, lsmp, n_rows, threads=cores, kind=kind)
    result = O.find_max_coa(m4, threads=cores, kind=kind)
    sec = time.perf_counter() - t_a
    del m4
    return result, sec, ("reference" if kind == "ref" else "port"), cores


def host_onsets(prob: syn.SynthProblem, position: str) -> np.ndarray:
    """Linear onsets on the host with the reference's own STA/LTA C code (CPU arm only)."""
    from oracle import qm_oracle as O

    kind = "ref" if O.have("ref") else "port"
    fn = O.overlapping_sta_lta if position == "classic" else O.centred_sta_lta
    return syn.onsets_from_waveforms(prob, lambda x, a, b: fn(x, a, b, kind=kind))


def reference_sample_size(cfg: syn.SynthConfig) -> int:
    """Output samples per reference step: the float64 volume stays near 6 GB, a step near 2-6 s."""
    return int(max(50, min(400, (6.5e9 / 8.0) // cfg.n_nodes)))


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cfg = global_config(syn.CONFIGS[args.workload], world, args.samples, args.weak)
    t_sub = reference_sample_size(cfg)
    # only the first fsmp + t_sub + lsmp samples are ever touched: the full-size geometry / LUT /
    # pads, and waveforms of just that length
    prob = syn.make_problem(cfg, with_waveforms=False)
    rng = np.random.default_rng(100 + cfg.seed)
    n_sta = len(prob.stations_xyz)
    prob.waveforms = rng.standard_normal((3 * n_sta, prob.fsmp + t_sub + prob.lsmp))
    onsets = host_onsets(prob, cfg.position)
    per_step = []
    kind = cores = None
    for i in range(args.warmup + args.steps):
        _, sec, kind, cores = cpu_reference_window(prob.traveltimes, onsets, prob.fsmp, prob.lsmp, cfg.n_rows, 0, t_sub)
        if i >= args.warmup:
            per_step.append(sec)
    sec = float(np.mean(per_step))
    value = cfg.n_nodes * t_sub / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg.name}: {cfg.grid[0]}x{cfg.grid[1]}x{cfg.grid[2]} grid, "
                               f"{cfg.n_rows} rows, bounded sample of {t_sub} of {cfg.n_samples} samples per step",
                   "n_nodes": cfg.n_nodes, "n_rows": cfg.n_rows, "n_samples": cfg.n_samples},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"full grid x first {t_sub} samples, migrate+find_max_coa (fresh zeroed volume per "
                                   f"step like core/lib.py:101), {args.steps} timed steps after {args.warmup} warm-up"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def preprocess_stage(eng, cfg, prob, args, peak, peak_src):
    """
    The upstream stage on the SAME workload: pre_process (stalta.py:203-209) of
    the raw trace block -- detrend, taper, zero-phase 2-corner Butterworth band-pass 2-16 Hz -- timed
    with CUDA events around its kernels, L2 flushed before every repetition, inputs resident in HBM.
    Algorithmic bytes: one float64 read and one float64 write per channel-sample (16 B); the
    implementation moves 40 B (moments pass 8 + two directions x [read 8 + write 8]: each direction is
    one kernel, tiles chained by a decoupled look-back).  The CPU figure is the oracle's scipy chain (what obspy runs for the reference:
    serial, per trace) on a bounded sample of traces.
    """
    from quakemigrate_b200 import preprocess as pp

    sr = cfg.sampling_rate
    sos = pp.butter_sos("bandpass", (2.0, 16.0), sr, 2)
    wf = prob.waveforms
    raw = eng.upload_waveforms(wf)
    out = eng.alloc_waveforms(*raw.shape)
    for _ in range(max(1, args.warmup)):
        eng.filter_waveforms(raw, sos, out=out)
    ms = []
    for _ in range(max(3, args.steps)):
        eng.flush_l2()
        _, t = eng.filter_waveforms(raw, sos, out=out, timing=True)
        ms.append(t)
    raw.free()
    out.free()
    t_ms = float(np.median(ms))
    n_cs = float(wf.shape[0]) * wf.shape[1]
    stage = {
        "what": "detrend + 5% cosine taper + zero-phase Butterworth band-pass [2, 16, 2] (pre_process)",
        "channels": int(wf.shape[0]), "samples_per_channel": int(wf.shape[1]), "dtype": "f64",
        "kernel_ms": t_ms, "value": n_cs / (t_ms * 1e-3), "unit": "channel-samples/s",
        "gpu_launches_per_call": 5,
        "roofline": {"bound": "hbm", "achieved": 16.0 * n_cs / (t_ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": 16.0 * n_cs / (t_ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_call": 16.0 * n_cs, "moved_bytes_model": 40.0 * n_cs,
                     "peak_source": peak_src},
    }
    if not args.no_cpu_baseline:
        from oracle import preprocess_oracle as PO  # CPU baseline leg only
        k = min(4, wf.shape[0])
        t0 = time.perf_counter()
        for c in range(k):
            PO.pre_process(wf[c], sr, [2.0, 16.0, 2], sos=sos)
        sec = time.perf_counter() - t0
        stage["cpu_baseline"] = {"value": k * wf.shape[1] / sec, "unit": "channel-samples/s", "cores": 1,
                                 "kind": "port",
                                 "sample": f"{k} of {wf.shape[0]} traces, scipy detrend + taper + sosfilt forward/backward "
                                           "(the calls obspy makes for the reference; serial per trace)"}
    return stage



def timestep_chain_stage(eng, cfg, prob, lut, S):
    """
    One detect timestep as QuakeScan._continuous_compute runs it (scan.py:434-462), from RAW traces
    in host memory to the integer channels ScanmSEED stores (io/scanmseed.py:102-130): H2D of the
    trace block, pre_process (stalta.py:203-209), STA/LTA + onset rows (stalta.py:364-470), the
    scan, the output stage, D2H of COA / COA_N / X / Y / Z.  Wall clock between device
    synchronisations; shows what the stages either side of the scan cost next to it.  Timing
    only: band-passing the synthetic traces changes the onset values, not the work.
    """
    from quakemigrate_b200 import preprocess as pp

    sos = pp.butter_sos("bandpass", (2.0, 16.0), cfg.sampling_rate, 2)
    h = cfg.node_spacing_km
    T = cfg.n_samples
    rows, parts = [], []
    for _ in range(3):
        eng.sync()
        t0 = time.perf_counter()
        raw = eng.upload_waveforms(prob.waveforms)
        eng.sync()
        ta = time.perf_counter()
        flt = eng.filter_waveforms(raw, sos)
        eng.sync()
        tb = time.perf_counter()
        dev = eng.onsets_from_device_waveforms(flt, prob.row_channels, prob.nsta, prob.nlta, position=cfg.position)
        eng.sync()
        t1 = time.perf_counter()
        parts.append((ta - t0, tb - ta, t1 - tb))
        eng.detect_device(lut, dev, prob.fsmp, prob.lsmp, S)
        eng.sync()
        t2 = time.perf_counter()
        out = eng.detect_fetch_quantised(T, cfg.grid, (0.0, 0.0, 0.0), (h, h, h), 1e3)
        t3 = time.perf_counter()
        rows.append((t1 - t0, t2 - t1, t3 - t2))
        raw.free(); flt.free(); dev.free()
    a = np.median(np.array(rows[1:]), axis=0) * 1e3
    b = np.median(np.array(parts[1:]), axis=0) * 1e3
    return {"what": "raw host traces -> H2D -> pre_process -> STA/LTA onset rows -> scan -> output stage -> host int32 channels",
            "before_scan_ms": float(a[0]),
            "before_scan_parts_ms": {"h2d_pageable": float(b[0]), "pre_process": float(b[1]), "onset_rows": float(b[2])},
            "scan_ms": float(a[1]), "output_stage_ms": float(a[2]),
            "total_ms": float(a.sum()), "h2d_bytes": int(prob.waveforms.nbytes),
            "d2h_bytes": int(sum(v.nbytes for v in out.values())),
            "note": "wall clock between device synchronisations, median of 2 after 1 warm-up, outside the timed region"}


# ------------------------------------------------------------------------------------------------
# parity of this run's result against the reference C, on random windows of the GLOBAL grid
# ------------------------------------------------------------------------------------------------
def parity_and_cpu_baseline(args, cfg, tt_global, prob, onsets_lin, res):
    """
    `res` = (max_coa, max_norm_coa, max_idx) of the e2e leg (the global result, identical on every
    rank).  Runs the reference on `--parity-windows` windows of `--parity-samples` output samples at
    seeded random offsets, compares, and times those runs (the CPU baseline of this line).
    """
    T, S = cfg.n_samples, cfg.n_rows
    n_win, t_sub = max(1, args.parity_windows), min(args.parity_samples, T)
    rng = np.random.default_rng(20260417)
    starts = sorted(int(v) for v in rng.integers(0, max(1, T - t_sub + 1), size=n_win))
    if n_win >= 2:  # always include the very first and the very last window (pads, tails)
        starts[0], starts[-1] = 0, T - t_sub
    worst = {"argmax_equal": True, "max_coa_max_rel": 0.0, "max_norm_coa_max_rel": 0.0}
    secs, kind, cores = [], None, None
    mismatches = 0
    for t0 in starts:
        ref, sec, kind, cores = cpu_reference_window(tt_global, onsets_lin, prob.fsmp, prob.lsmp, S, t0, t_sub)
        secs.append(sec)
        sl = slice(t0, t0 + t_sub)
        eq = np.array_equal(res[2][sl], ref[2])
        mismatches += int(np.sum(res[2][sl] != ref[2]))
        worst["argmax_equal"] = bool(worst["argmax_equal"] and eq)
        worst["max_coa_max_rel"] = max(worst["max_coa_max_rel"], float(np.max(np.abs(res[0][sl] - ref[0]) / ref[0])))
        worst["max_norm_coa_max_rel"] = max(worst["max_norm_coa_max_rel"],
                                            float(np.max(np.abs(res[1][sl] - ref[1]) / ref[1])))
    sec = float(np.median(secs))
    parity = {"windows": [[t0, t0 + t_sub] for t0 in starts], "samples": n_win * t_sub,
              "argmax_mismatches": mismatches, **worst,
              "tolerance": {"argmax": "bit-exact", "max_coa": 1e-5, "max_norm_coa": 1e-5},
              "pass": bool(worst["argmax_equal"] and worst["max_coa_max_rel"] <= 1e-5 and
                           worst["max_norm_coa_max_rel"] <= 1e-5),
              "note": "e2e result of this run (global, all ranks) against the reference C on the full GLOBAL grid"}
    cpu = {"value": cfg.n_nodes * t_sub / sec, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"full {cfg.grid[0]}x{cfg.grid[1]}x{cfg.grid[2]} grid x {t_sub}-sample windows at {starts} of {T} "
                     f"samples, migrate+find_max_coa, median of {n_win} ({sec:.2f} s each)"}
    return parity, cpu


# ------------------------------------------------------------------------------------------------
# GPU arm: detect
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int, dist):
    from quakemigrate_b200.core import Engine
    from quakemigrate_b200.sharding import balanced_range, slab_bounds

    base = syn.CONFIGS[args.workload]
    cfg = global_config(base, world, args.samples, args.weak)
    nx, ny, nz = cfg.grid
    plane = ny * nz
    if args.weak:
        x0, x1 = rank * base.grid[0], (rank + 1) * base.grid[0]
        node_first, N_local = x0 * plane, (x1 - x0) * plane
    elif args.plane_slabs:
        x0, x1 = slab_bounds(nx, world, rank)
        node_first, N_local = x0 * plane, (x1 - x0) * plane
    else:
        # equal shares of the C-order node list in whole z-columns: the cuts may fall inside an x-plane
        node_first, N_local = balanced_range(nx, ny, nz, world, rank)
        x0, x1 = node_first // plane, -(-(node_first + N_local) // plane)
    prob = syn.make_problem(cfg, with_waveforms=True, x_range=(x0, x1))
    N_global, T, S = cfg.n_nodes, cfg.n_samples, cfg.n_rows
    P = prob.n_total

    eng = Engine(local_rank, variant=args.variant if args.variant >= 0 else None, precision=args.precision,
                 qgroup=args.qgroup)
    if args.pre_only:
        print(json.dumps(preprocess_stage(eng, base, prob, args, *measured_peak_hbm())), flush=True)
        eng.close()
        return
    if world > 1:
        ids = [Engine.new_comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, group=CTRL_GROUP)
        eng.init_comm(ids[0], world, rank)

    rows = prob.traveltimes.reshape(-1, S)[node_first - x0 * plane: node_first - x0 * plane + N_local]
    lut = eng.upload_lut_range(rows, ny, nz, node_first, N_global)
    del rows
    dev_on, onsets_lin = eng.onsets_from_waveforms(prob.waveforms, prob.row_channels, prob.nsta, prob.nlta,
                                                   position=cfg.position, return_onsets=True)
    info = lut.info()

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing: `value` ---------------------------------------------------
    for _ in range(args.warmup):
        eng.detect_device(lut, dev_on, prob.fsmp, prob.lsmp, S)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = eng.launch_count
    scan_ms, total_ms, tail_ms, slow, syncs = [], [], [], 0, 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.flush_l2()
        ms = eng.detect_device(lut, dev_on, prob.fsmp, prob.lsmp, S)
        scan_ms.append(ms["scan_ms"])
        total_ms.append(ms["total_ms"])
        tail_ms.append(ms["finalize_ms"])
        slow = max(slow, ms["slow_samples"])
        syncs = max(syncs, eng.last_scan_info["host_syncs_in_timed_region"])
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0  # own kernels only (L2 flushes are memsets)
    clocks = sampler.stop()
    dev_total_ms = float(np.sum(total_ms))
    scan_info = eng.last_scan_info

    # ---- end-to-end through the host-buffer call: `e2e` ----------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    Engine.pin(onsets_lin)
    eng.detect(lut, onsets_lin, prob.fsmp, prob.lsmp, S)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = eng.detect(lut, onsets_lin, prob.fsmp, prob.lsmp, S)
    barrier()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    Engine.unpin(onsets_lin)

    # ---- max over ranks ----------------------------------------------------------------------
    if world > 1:
        import torch
        t = torch.tensor([dev_total_ms, e2e_sec, float(np.mean(scan_ms)), wall, float(np.mean(tail_ms)), float(slow)],
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=CTRL_GROUP)
        dev_total_ms, e2e_sec, scan_mean, wall, tail_mean, slow = [float(x) for x in t]
        slow = int(slow)
        # per-rank means: a rank's tail includes its wait for the slowest rank's scan at the exchange
        mine = torch.tensor([float(np.mean(scan_ms)), float(np.mean(tail_ms))], dtype=torch.float64)
        every = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(every, mine, group=CTRL_GROUP)
        per_rank = {"scan_ms": [round(float(x[0]), 2) for x in every], "tail_ms": [round(float(x[1]), 2) for x in every]}
    else:
        scan_mean, tail_mean = float(np.mean(scan_ms)), float(np.mean(tail_ms))
        per_rank = None
    ms_per_step = dev_total_ms / args.steps
    value = N_global * T / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (k_scan): SURVEY.md section 8(d) -----------------------
    peak, peak_src = measured_peak_hbm()
    b_alg = 4.0 * N_local * T * (S + 1) + 4.0 * N_local * S + 4.0 * S * P
    achieved = b_alg / (scan_mean * 1e-3) / 1e9
    bpg = eng.last_scan_bytes_per_gather
    traffic, traffic_src = (None, None)
    if world == 1 and not args.samples:
        traffic, traffic_src = ncu_traffic(base.name, info["variant"], scan_info["operand_bits"])
    measured_pipe = eng.smem_pipe_peak()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    ga = float(N_local) * T * S / (scan_mean * 1e-3)
    smem_peak = 148 * 32 * sm_mhz * 1e6
    # shared-memory bytes one gathered operand costs: 4 (float32 rows) or 2 (16-bit fixed-point
    # rows, two samples per LDS.32); the pipe moves 128 B/clk/SM either way

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{cfg.name}: {cfg.grid[0]}x{cfg.grid[1]}x{cfg.grid[2]} grid "
                        f"(nodes [{node_first}, {node_first + N_local}) on rank {rank}), {S} onset rows, "
                        f"{T} samples @ {cfg.sampling_rate} Hz, {cfg.position} STA/LTA, detect() scan",
            "n_nodes": N_global, "n_rows": S, "n_samples": T, "fsmp": prob.fsmp, "lsmp": prob.lsmp,
            "parallelism": f"node-slab x{world} (" + ("one workload-sized slab per GPU" if args.weak else
                           "fixed grid, x-plane slabs" if args.plane_slabs else
                           "fixed grid, equal contiguous node ranges in whole z-columns") + ")",
            "cache": "L2 flushed (256 MiB memset) before every timed step",
            "nodes_per_tile": info["nodes_per_tile"], "max_tile_span": info["max_tile_span"],
            "variant": info["variant"], "duplicate_fraction": round(info["duplicate_fraction"], 4),
            "scan_operands": f"{scan_info['operand_bits']}-bit fixed point (2 per LDS.32), node sums verified "
                             "against float32 windows" if bpg == 2 else "float32",
            "timing": "sum of per-step CUDA-event times on the launching stream (scan + merge + exchange + "
                      "candidates + slow path + verification windows), max over ranks",
            "git": git_head(),
        },
        "e2e": {"value": N_global * T / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(onsets_lin.nbytes),
                "d2h_bytes_per_step": int(3 * 8 * T), "steps": e2e_steps,
                "note": "Engine.detect(): pinned host float64 onsets -> H2D -> clip/log/quantise -> scan -> reduce -> "
                        "D2H of the three traces, wall clock between device synchronisations and rank barriers; "
                        "travel-time table resident (uploaded once per availability pattern); loop capped at "
                        "--e2e-steps"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_scan", "kernel_ms": scan_mean,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg,
            "note": "effective bandwidth per SURVEY 8(d): operands are served from shared memory after "
                    "one compulsory HBM pass, so frac may exceed 1; the binding resources are the "
                    "shared-memory pipe and instruction issue (smem_pipe below)",
            "smem_pipe": {"achieved_gather_adds_per_s": ga, "bytes_per_gather": bpg,
                          "peak_gather_adds_per_s": smem_peak * 4 / bpg,
                          "peak_gather_adds_per_s_fp32_operands": smem_peak,
                          "frac": ga * bpg / (smem_peak * 4), "sm_mhz_used": sm_mhz,
                          "measured_peak_gather_adds_per_s": measured_pipe * 4 / bpg,
                          "frac_of_measured": ga * bpg / (measured_pipe * 4),
                          "measured_how": "qmb_microbench_smem_pipe: bare LDS.32+FADD streams, 12 warps/SM, same box, same run"},
        },
        "wall_s_timed_region": wall,
        "slow_path_samples": slow,
        "tail_ms_per_step": tail_mean,
        "per_rank": per_rank,
        "scan": {"operand_bytes": scan_info["operand_bytes"], "operand_bits": scan_info["operand_bits"],
                 "fell_back_fp32": scan_info["fell_back_fp32"],
                 "check_windows": scan_info["check_windows"], "check_samples": scan_info["check_samples"],
                 "check_max_rel_err": scan_info["check_max_rel_err"], "check_ms": scan_info["check_ms"],
                 "two_candidate_samples": scan_info["two_candidate_samples"],
                 "host_syncs_in_timed_region": syncs},
    }

    if rank == 0 and world == 1:
        line["stages"] = {"preprocess": preprocess_stage(eng, cfg, prob, args, peak, peak_src),
                          "timestep_chain": timestep_chain_stage(eng, cfg, prob, lut, S)}
    if rank == 0 and not args.no_cpu_baseline:
        # the GLOBAL table on rank 0's host (the reference needs all nodes)
        tt_global = prob.traveltimes if (x0 == 0 and x1 == nx) else syn.make_traveltimes(cfg)[0]
        line["parity_vs_reference_sample"], line["cpu_baseline"] = parity_and_cpu_baseline(
            args, cfg, tt_global, prob, onsets_lin, res)
        del tt_global
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    del res
    eng.close()


# ------------------------------------------------------------------------------------------------
# GPU arm: locate (BASELINE configs[3]): span scan -> marginal window -> 3-D map stage
# ------------------------------------------------------------------------------------------------
def run_locate(args, rank: int, world: int, local_rank: int, dist):
    """
    One step = one located event (QuakeScan._locate_events body, scan.py:508-531): the max traces
    over the whole span (16-bit rows), otime -> marginal window (io/event.py:421-439), the window
    re-migrated with float32 rows and summed over time on the device (scan.py:720), and
    _calculate_location on the device-resident map (scan.py:696-733).  With N GPUs every rank
    locates its own replica of the event ("replicas only": events are independent).
    """
    from quakemigrate_b200.core import Engine
    from quakemigrate_b200.locate import calculate_location_device
    from quakemigrate_b200.lut import Grid3D
    from quakemigrate_b200.scan import marginal_window_indices

    cfg = syn.CONFIGS[args.workload].scaled(n_samples=args.samples or None, name=args.workload)
    prob = syn.make_problem(cfg, with_waveforms=True)
    N, T, S = cfg.n_nodes, cfg.n_samples, cfg.n_rows
    mw = 2.0  # QuakeScan.marginal_window default (scan.py:236)
    eng = Engine(local_rank, variant=args.variant if args.variant >= 0 else None, precision=args.precision,
                 qgroup=args.qgroup)
    lut = eng.upload_lut(prob.traveltimes)
    dev_on, onsets_lin = eng.onsets_from_waveforms(prob.waveforms, prob.row_channels, prob.nsta, prob.nlta,
                                                   position=cfg.position, return_onsets=True)
    grid = Grid3D(cfg.grid, [cfg.node_spacing_km] * 3, [0.0, 0.0, 0.0])
    info = lut.info()

    def event(dev, stamps=None):
        """One located event; the CUDA-event timer (started by the caller) stops after the last device op."""
        ms = eng.detect_device(lut, dev, prob.fsmp, prob.lsmp, S)
        mx, nm, ix = eng.detect_fetch(T)
        i0, i1, imax = marginal_window_indices(mx, mw, cfg.sampling_rate)
        if stamps is not None:
            eng.sync(); stamps.append(time.perf_counter())
        dmap = eng.locate_window_device(lut, dev, prob.fsmp, prob.lsmp, S, i0, i1)
        if stamps is not None:
            eng.sync(); stamps.append(time.perf_counter())
        box = {}

        def device_done():
            box["ms"] = eng.timer_stop()
            if stamps is not None:
                stamps.append(time.perf_counter())
        loc = calculate_location_device(dmap, grid, engine_timer=device_done)
        dmap.free()
        return {"scan": ms, "window": (i0, i1), "peak_sample": imax, "idx": ix[imax], "device_ms": box["ms"]}, loc

    for _ in range(max(1, args.warmup)):
        eng.timer_start()
        event(dev_on)
    sampler = ClockSampler(local_rank)
    eng.sync()
    if world > 1:
        dist.barrier()
    sampler.start()
    launches0 = eng.launch_count
    dev_ms, scan_ms, span_total_ms = [], [], []
    for _ in range(args.steps):
        eng.flush_l2()
        eng.timer_start()
        out, loc = event(dev_on)
        dev_ms.append(out["device_ms"])
        scan_ms.append(out["scan"]["scan_ms"])
        span_total_ms.append(out["scan"]["total_ms"])
    eng.sync()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count - launches0
    clocks = sampler.stop()
    # stage breakdown (wall clock between device synchronisations; outside the timed region)
    stage = []
    for _ in range(3):
        stamps = [time.perf_counter()]
        eng.timer_start()
        event(dev_on, stamps)
        stamps.append(time.perf_counter())
        stage.append(np.diff(stamps))

    # ---- e2e: host onsets in, location out (H2D of the onsets, host fits included) ---------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    Engine.pin(onsets_lin)
    t_e2e, t_up = [], []
    for i in range(e2e_steps + 1):
        eng.sync()
        t0 = time.perf_counter()
        d = eng.upload_onsets(onsets_lin)
        eng.sync()
        t1 = time.perf_counter()
        eng.timer_start()
        out, loc = event(d)
        d.free()
        eng.sync()
        if i:
            t_e2e.append(time.perf_counter() - t0)
            t_up.append(t1 - t0)
    Engine.unpin(onsets_lin)
    e2e_sec = float(np.mean(t_e2e))
    dev_total_ms = float(np.sum(dev_ms))
    scan_mean = float(np.mean(scan_ms))
    if world > 1:
        import torch
        t = torch.tensor([dev_total_ms, e2e_sec, scan_mean], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=CTRL_GROUP)
        dev_total_ms, e2e_sec, scan_mean = [float(x) for x in t]
    ms_per_step = dev_total_ms / args.steps
    value = world * N * T / (ms_per_step * 1e-3)
    st = np.median(np.array(stage), axis=0) * 1e3  # span pass, window pass, device map stage, host fits (ms, wall)
    i0, i1 = out["window"]
    Wn = i1 - i0

    peak, peak_src = measured_peak_hbm()
    P = prob.n_total
    b_alg = 4.0 * N * T * (S + 1) + 4.0 * N * S + 4.0 * S * P
    achieved = b_alg / (scan_mean * 1e-3) / 1e9
    bpg = eng.last_scan_bytes_per_gather
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    smem_peak = 148 * 32 * sm_mhz * 1e6
    ga = float(N) * T * S / (scan_mean * 1e-3)
    win_hbm = 8.0 * N + 2.0 * N * S + 4.0 * S * (Wn + prob.lsmp)  # map written + tau16 tiles read + row windows
    planted = np.unravel_index(prob.events[0][0], cfg.grid)
    found = np.unravel_index(int(out["idx"]), cfg.grid)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{cfg.name} locate: {cfg.grid[0]}x{cfg.grid[1]}x{cfg.grid[2]} grid, {S} onset rows, "
                        f"{T} samples @ {cfg.sampling_rate} Hz, {cfg.position} STA/LTA, marginal window "
                        f"{Wn} samples, Gaussian / covariance / spline location fits",
            "mode": "locate", "n_nodes": N, "n_rows": S, "n_samples": T, "window_samples": Wn,
            "parallelism": "one event per GPU (replicas only)" if world > 1 else "1 GPU",
            "cache": "L2 flushed (256 MiB memset) before every timed step",
            "variant": info["variant"], "timing": "CUDA events on the launching stream from the span scan to the "
            "last device operation of the event (host gaps between them included), max over ranks",
            "git": git_head(),
        },
        "e2e": {"value": world * N * T / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(onsets_lin.nbytes),
                "d2h_bytes_per_step": int(3 * 8 * T + loc["pcie_bytes"]), "steps": e2e_steps,
                "seconds_per_event": e2e_sec,
                "note": "pinned host onsets -> H2D -> span scan -> D2H of the traces -> window pass -> device map "
                        "stage -> windows/moments D2H -> host fits (RBF spline, pinv)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel": "k_scan (span pass)", "kernel_ms": scan_mean, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": b_alg,
            "note": "effective bandwidth per SURVEY 8(d), see the detect line; shared-memory pipe fraction below",
            "smem_pipe": {"achieved_gather_adds_per_s": ga, "bytes_per_gather": bpg,
                          "frac": ga * bpg / (smem_peak * 4), "sm_mhz_used": sm_mhz},
        },
        "stages": {
            "e2e_steps_s": [round(float(v), 4) for v in t_e2e],
            "e2e_onset_upload_and_row_preparation_s": [round(float(v), 4) for v in t_up],
            "timed_steps_device_ms": [round(float(v), 2) for v in dev_ms],
            "timed_steps_span_scan_ms": [round(float(v), 2) for v in span_total_ms],
            "span_pass_ms": float(st[0]), "window_pass_ms": float(st[1]), "device_map_stage_ms": float(st[2]),
            "host_fits_ms": float(st[3]),
            "note": "wall clock between device synchronisations, measured outside the timed region",
            "window_pass": {
                "what": "k_scan MODE 1 on float32 rows: exp(stack/available) summed over the window per node",
                "roofline": {"bound": "hbm", "achieved": win_hbm / (st[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": win_hbm / (st[1] * 1e-3) / 1e9 / peak, "true_hbm_bytes_model": win_hbm,
                             "note": "true HBM bytes (8 N written, table and row windows read once); the pass is "
                                     "bound by the shared-memory pipe like the scan: gather-adds/s below"},
                "gather_adds_per_s": float(N) * Wn * S / (st[1] * 1e-3)},
            "location_pcie_bytes_per_event": loc["pcie_bytes"],
        },
        "result": {"peak_node": [int(v) for v in found], "planted_node": [int(v) for v in planted],
                   "peak_sample": int(out["peak_sample"]), "planted_origin_sample": int(prob.events[0][1]),
                   "gaussian_location_km": [float(v) for v in loc["gaussian"][0]],
                   "gaussian_uncertainty_km": [float(v) for v in loc["gaussian"][1]],
                   "covariance_location_km": [float(v) for v in loc["covariance"][0]],
                   "spline_location_km": [float(v) for v in loc["spline"]]},
        "slow_path_samples": int(out["scan"]["slow_samples"]),
    }
    if rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"], line["parity_vs_reference_sample"] = locate_cpu_baseline(
            args, cfg, prob, onsets_lin, eng, lut, dev_on, i0, i1, out)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()


def locate_cpu_baseline(args, cfg, prob, onsets_lin, eng, lut, dev_on, i0, i1, out):
    """
    The reference's locate arithmetic on a bounded sample, timed on the host cores: migrate +
    find_max_coa on 100 span samples (the per-sample cost of the span pass), migrate on the first
    64 samples of the marginal window + np.sum(axis=-1) (scan.py:720), and _gaufilt3d
    (scan.py:1008-1043, oracle restatement) on the full map.  Extrapolated linearly to one event.
    The window leg doubles as the parity check of the device's marginal map.
    """
    from oracle import qm_oracle as O

    N, T, S = cfg.n_nodes, cfg.n_samples, cfg.n_rows
    t_span = min(100, T)
    ref, sec_span, kind, cores = cpu_reference_window(prob.traveltimes, onsets_lin, prob.fsmp, prob.lsmp, S, 0, t_span)
    w_sub = min(64, i1 - i0)
    k = "ref" if O.have("ref") else "port"
    sub = np.ascontiguousarray(onsets_lin[:, i0: i0 + prob.fsmp + w_sub + prob.lsmp])
    t_a = time.perf_counter()
    m4 = O.migrate(sub, prob.traveltimes, prob.fsmp, prob.lsmp, S, threads=cores, kind=k)
    coa_ref = m4.sum(axis=-1)
    sec_win = time.perf_counter() - t_a
    del m4
    got = eng.locate_window(lut, dev_on, prob.fsmp, prob.lsmp, S, i0, i0 + w_sub)
    rel = float(np.max(np.abs(got - coa_ref) / coa_ref))
    t_a = time.perf_counter()
    O.np_gaufilt3d(coa_ref / coa_ref.max())
    sec_gau = time.perf_counter() - t_a
    res = eng.detect_fetch(T)
    est = sec_span / t_span * T + sec_win / w_sub * (i1 - i0) + sec_gau
    cpu = {"value": N * T / est, "unit": UNIT, "cores": cores, "kind": kind,
           "seconds_per_event_extrapolated": est,
           "sample": f"span: full grid x {t_span} of {T} samples ({sec_span:.2f} s); window: full grid x {w_sub} of "
                     f"{i1 - i0} samples + sum over time ({sec_win:.2f} s); _gaufilt3d of the full map, oracle "
                     f"restatement ({sec_gau:.2f} s); extrapolated linearly in samples"}
    parity = {"span_samples": t_span, "argmax_equal": bool(np.array_equal(res[2][:t_span], ref[2])),
              "max_coa_max_rel": float(np.max(np.abs(res[0][:t_span] - ref[0]) / ref[0])),
              "max_norm_coa_max_rel": float(np.max(np.abs(res[1][:t_span] - ref[1]) / ref[1])),
              "window_samples": w_sub, "coa_map_max_rel": rel,
              "tolerance": {"argmax": "bit-exact", "values": 1e-5}}
    parity["pass"] = bool(parity["argmax_equal"] and parity["max_coa_max_rel"] <= 1e-5 and
                          parity["max_norm_coa_max_rel"] <= 1e-5 and rel <= 1e-5)
    return cpu, parity


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        dist = dist_mod
        global CTRL_GROUP
        CTRL_GROUP = dist.new_group(backend="gloo")  # control plane: ids, timings (CPU tensors)
    elif args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (see the module docstring)")
    if args.mode == "locate":
        run_locate(args, rank, world, local_rank, dist)
    else:
        run_ours(args, rank, world, local_rank, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
