# -*- coding: utf-8 -*-
"""
Benchmark of the coalescence hot path (fused migrate + stack + max-reduce) on B200.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own OpenMP C path on host cores

One "step" = one pass of the hot path over the workload: for every one of the N*T
(node, sample) pairs stack S onset samples, exponentiate, and reduce over nodes to the per-sample
(max, arg-max, normalised max) -- what QuakeScan._compute does per detect timestep
(quakemigrate/signal/scan.py:635-638).  Metric: node-samples / second (whole job).

Workload: BASELINE.json configs[1] (101x101x51 grid, 32 onset rows, 120000 samples @ 100 Hz,
classic STA/LTA) per GPU; with N GPUs the grid is N such slabs stacked along x and sharded by
x-slab (weak scaling), every rank ends with the global result after one NCCL all-gather of the
per-sample candidates.

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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(syn.CONFIGS))
    ap.add_argument("--variant", type=int, default=-1, help="kernel tiling variant id; -1 = library default")
    ap.add_argument("--samples", type=int, default=0, help="override n_samples (debug)")
    ap.add_argument("--precision", default=None, choices=["auto", "q16", "fp32"],
                    help="operand format of the scan's gathers (library default: auto -> q16 at bench sizes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pre-only", action="store_true",
                    help="profiling aid: run only the pre-processing stage on the workload's full-length traces")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: shard the FIXED workload grid over the ranks (default: weak, one "
                         "workload-sized slab per GPU)")
    return ap.parse_args()


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy burst)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def global_config(base: syn.SynthConfig, world: int, n_samples: int) -> syn.SynthConfig:
    nx, ny, nz = base.grid
    return base.scaled(grid=(nx * world, ny, nz), n_samples=n_samples or base.n_samples,
                       name=f"{base.name}x{world}")


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
def cpu_reference_sample(cfg: syn.SynthConfig, prob: syn.SynthProblem, onsets_lin: np.ndarray,
                         t_sub: int, warm: int, reps: int, return_result: bool = False):
    """
    Time migrate + find_max_coa of the reference on a bounded sample: the full grid, the first
    `t_sub` output samples.  Returns (node-samples/s, kind, cores, seconds per rep).
    """
    from oracle import qm_oracle as O  # the ONE place bench.py touches oracle/: the CPU baseline

    kind = "ref" if O.have("ref") else "port"
    cores = O.default_threads()
    t_sub = min(t_sub, cfg.n_samples)
    sub = np.ascontiguousarray(onsets_lin[:, : prob.fsmp + t_sub + prob.lsmp])
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        m4 = O.migrate(sub, prob.traveltimes, prob.fsmp, prob.lsmp, cfg.n_rows, threads=cores, kind=kind)
        result = O.find_max_coa(m4, threads=cores, kind=kind)
        dt = time.perf_counter() - t0
        del m4
        if i >= warm:
            times.append(dt)
    sec = float(np.median(times))
    out = (cfg.n_nodes * t_sub / sec, ("reference" if kind == "ref" else "port"), cores, sec)
    return out + (result,) if return_result else out


def host_onsets(prob: syn.SynthProblem, position: str) -> np.ndarray:
    """Linear onsets on the host with the reference's own STA/LTA C code (CPU arm only)."""
    from oracle import qm_oracle as O

    kind = "ref" if O.have("ref") else "port"
    fn = O.overlapping_sta_lta if position == "classic" else O.centred_sta_lta
    return syn.onsets_from_waveforms(prob, lambda x, a, b: fn(x, a, b, kind=kind))


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cfg = global_config(syn.CONFIGS[args.workload], world, args.samples)
    t_sub = 400
    # only the first fsmp + t_sub + lsmp samples are ever touched: build a short problem with the
    # same geometry / LUT / pads
    prob = syn.make_problem(cfg, with_waveforms=False)
    short = cfg.scaled(n_samples=t_sub, name=cfg.name)
    rng = np.random.default_rng(100 + cfg.seed)
    n_sta = len(prob.stations_xyz)
    prob.waveforms = rng.standard_normal((3 * n_sta, prob.fsmp + t_sub + prob.lsmp))
    onsets = host_onsets(prob, cfg.position)
    per_step = []
    kind = cores = None
    for i in range(args.warmup + args.steps):
        v, kind, cores, sec = cpu_reference_sample(short, prob, onsets, t_sub, 0, 1)
        if i >= args.warmup:
            per_step.append(sec)
    sec = float(np.mean(per_step))
    value = cfg.n_nodes * t_sub / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{cfg.name}: {cfg.grid[0]}x{cfg.grid[1]}x{cfg.grid[2]} grid, "
                               f"{cfg.n_rows} rows, bounded sample of {t_sub} of {cfg.n_samples} samples per step",
                   "n_nodes": cfg.n_nodes, "n_rows": cfg.n_rows, "n_samples": cfg.n_samples},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"full grid x first {t_sub} samples, migrate+find_max_coa, "
                                   f"{args.steps} timed steps after {args.warmup} warm-up"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def preprocess_stage(eng, cfg, prob, args, peak, peak_src):
    """
    The upstream stage on the SAME workload: pre_process (stalta.py:203-209) of
    the raw trace block -- detrend, taper, zero-phase 2-corner Butterworth band-pass 2-16 Hz -- timed
    with CUDA events around its kernels, L2 flushed before every repetition, inputs resident in HBM.
    Algorithmic bytes: one float64 read and one float64 write per channel-sample (16 B); the
    implementation moves 56 B (moments pass + two directions x [state pass read, apply pass
    read + write]).  The CPU figure is the oracle's scipy chain (what obspy runs for the reference:
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
        "gpu_launches_per_call": 8,
        "roofline": {"bound": "hbm", "achieved": 16.0 * n_cs / (t_ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": 16.0 * n_cs / (t_ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_call": 16.0 * n_cs, "moved_bytes_model": 56.0 * n_cs,
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


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int, dist):
    from quakemigrate_b200.core import Engine

    base = syn.CONFIGS[args.workload]
    if args.strong:
        from quakemigrate_b200.sharding import slab_bounds
        cfg = base.scaled(n_samples=args.samples or base.n_samples, name=base.name)
        x0, x1 = slab_bounds(cfg.grid[0], world, rank)
    else:
        cfg = global_config(base, world, args.samples)
        x0, x1 = rank * base.grid[0], (rank + 1) * base.grid[0]
    prob = syn.make_problem(cfg, with_waveforms=True, x_range=(x0, x1))
    N_local = (x1 - x0) * cfg.grid[1] * cfg.grid[2]
    N_global, T, S = cfg.n_nodes, cfg.n_samples, cfg.n_rows
    P = prob.n_total

    eng = Engine(local_rank, variant=args.variant if args.variant >= 0 else None, precision=args.precision)
    if args.pre_only:
        full = syn.make_problem(global_config(base, 1, 0), with_waveforms=True)
        print(json.dumps(preprocess_stage(eng, base, full, args, *measured_peak_hbm())), flush=True)
        eng.close()
        return
    if world > 1:
        ids = [Engine.new_comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, group=CTRL_GROUP)
        eng.init_comm(ids[0], world, rank)

    lut = eng.upload_lut(prob.traveltimes, node_offset=x0 * cfg.grid[1] * cfg.grid[2], n_nodes_global=N_global)
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
    scan_ms, total_ms, slow = [], [], 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.flush_l2()
        ms = eng.detect_device(lut, dev_on, prob.fsmp, prob.lsmp, S)
        scan_ms.append(ms["scan_ms"])
        total_ms.append(ms["total_ms"])
        slow = max(slow, ms["slow_samples"])
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0  # own kernels only (L2 flushes are memsets)
    clocks = sampler.stop()
    dev_total_ms = float(np.sum(total_ms))

    # ---- end-to-end through the host-buffer call: `e2e` ----------------------------------------
    Engine.pin(onsets_lin)
    for _ in range(2):
        eng.detect(lut, onsets_lin, prob.fsmp, prob.lsmp, S)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = eng.detect(lut, onsets_lin, prob.fsmp, prob.lsmp, S)
    barrier()
    e2e_sec = (time.perf_counter() - t0) / args.steps
    Engine.unpin(onsets_lin)

    # ---- max over ranks ----------------------------------------------------------------------
    if world > 1:
        import torch
        t = torch.tensor([dev_total_ms, e2e_sec, float(np.mean(scan_ms)), wall], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=CTRL_GROUP)
        dev_total_ms, e2e_sec, scan_mean, wall = [float(x) for x in t]
    else:
        scan_mean = float(np.mean(scan_ms))
    ms_per_step = dev_total_ms / args.steps
    value = N_global * T / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (k_scan): SURVEY.md section 8(d) -----------------------
    peak, peak_src = measured_peak_hbm()
    b_alg = 4.0 * N_local * T * (S + 1) + 4.0 * N_local * S + 4.0 * S * P
    achieved = b_alg / (scan_mean * 1e-3) / 1e9
    traffic = None
    prof = ROOT / "profiles" / "scan_ncu_summary.json"
    if prof.exists():
        try:
            pj = json.loads(prof.read_text())
            if pj.get("workload") == base.name:
                traffic = pj.get("dram_bytes_per_launch")
        except Exception:
            pass
    measured_pipe = eng.smem_pipe_peak()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    ga = float(N_local) * T * S / (scan_mean * 1e-3)
    smem_peak = 148 * 32 * sm_mhz * 1e6
    # shared-memory bytes one gathered operand costs: 4 (float32 rows) or 2 (16-bit fixed-point
    # rows, two samples per LDS.32); the pipe moves 128 B/clk/SM either way
    bpg = eng.last_scan_bytes_per_gather

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{cfg.name}: {cfg.grid[0]}x{cfg.grid[1]}x{cfg.grid[2]} grid "
                        f"({x1 - x0}x{cfg.grid[1]}x{cfg.grid[2]} x-slab on rank {rank}), {S} onset rows, "
                        f"{T} samples @ {cfg.sampling_rate} Hz, {cfg.position} STA/LTA, detect() scan",
            "n_nodes": N_global, "n_rows": S, "n_samples": T, "fsmp": prob.fsmp, "lsmp": prob.lsmp,
            "parallelism": f"node-slab x{world}", "cache": "L2 flushed (256 MiB memset) before every timed step",
            "nodes_per_tile": info["nodes_per_tile"], "max_tile_span": info["max_tile_span"],
            "variant": info["variant"], "duplicate_fraction": round(info["duplicate_fraction"], 4),
            "scan_operands": "16-bit fixed point (2 per LDS.32)" if bpg == 2 else "float32",
            "timing": "sum of per-step CUDA-event times on the launching stream, max over ranks",
        },
        "e2e": {"value": N_global * T / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(onsets_lin.nbytes),
                "d2h_bytes_per_step": int(3 * 8 * T),
                "note": "Engine.detect(): host float64 onsets -> H2D -> clip/log -> scan -> reduce -> D2H; "
                        "travel-time table resident (uploaded once per availability pattern)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": "k_scan", "kernel_ms": scan_mean, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": b_alg,
            "note": "effective bandwidth per SURVEY 8(d): operands are served from shared memory after "
                    "one compulsory HBM pass, so frac may exceed 1; the binding pipe is shared memory",
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
    }

    if rank == 0 and world == 1:
        line["stages"] = {"preprocess": preprocess_stage(eng, cfg, prob, args, peak, peak_src)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, kind, cores, sec, ref = cpu_reference_sample(cfg, prob, onsets_lin, 600, 1, 2, return_result=True)
        n_cmp = len(ref[0])
        line["parity_vs_reference_sample"] = {
            "samples": n_cmp, "argmax_equal": bool(np.array_equal(res[2][:n_cmp], ref[2])),
            "max_coa_max_rel": float(np.max(np.abs(res[0][:n_cmp] - ref[0]) / ref[0])),
            "max_norm_coa_max_rel": float(np.max(np.abs(res[1][:n_cmp] - ref[1]) / ref[1])),
            "note": "e2e result of this run against the reference C on the CPU-baseline sample"}
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"full {cfg.grid[0]}x{cfg.grid[1]}x{cfg.grid[2]} grid x first 600 of {T} "
                                          f"samples, migrate+find_max_coa, median of 2 after 1 warm-up ({sec:.2f} s each)"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    del res
    eng.close()


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
    run_ours(args, rank, world, local_rank, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
