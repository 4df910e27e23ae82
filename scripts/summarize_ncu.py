#!/usr/bin/env python
"""
Condense ncu artefacts brought back in gpurun_out/ into small tracked files under profiles/.

    python scripts/summarize_ncu.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof_scan.ncu-rep]
                                          [--workload cfg2] [--skip-launches N]
"""
import argparse
import csv
import io
import json
import pathlib
import re
import subprocess
from collections import defaultdict

ROOT = pathlib.Path(__file__).resolve().parents[1]

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def short(name: str) -> str:
    m = re.search(r"(k_[a-z_0-9]+)(?:ILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)E)?", name)
    if not m:
        m2 = re.search(r"(k_[a-z_0-9]+)", name)
        return m2.group(1) if m2 else name[:60]
    return m.group(1) + ("<%s,%s,%s,%s>" % m.groups()[1:] if m.group(2) else "")


def launches(path, skip):
    text = pathlib.Path(path).read_text()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = defaultdict(lambda: [0, 0.0])
    seq = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        seq.append((short(r["Kernel Name"]), ns))
    for name, ns in seq[skip:]:
        per[name][0] += 1
        per[name][1] += ns
    total = sum(v[1] for v in per.values())
    table = sorted(((k, v[0], v[1] / 1e6, v[1] / total) for k, v in per.items()), key=lambda x: -x[2])
    return table, total / 1e6, len(seq)


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h == "Kernel Name":
            d[h] = (v, u)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--skip-launches", type=int, default=0)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    prof = ROOT / "profiles"
    prof.mkdir(exist_ok=True)
    lines = [f"# {a.tag}", "", a.note, ""]
    if a.launches:
        table, total, n = launches(a.launches, a.skip_launches)
        lines += [f"## launch list ({pathlib.Path(a.launches).name}; {n} launches captured, first {a.skip_launches} skipped)",
                  "", "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)", "",
                  "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
        lines += [f"| {k} | {c} | {ms:.3f} | {sh * 100:.2f}% |" for k, c, ms, sh in table]
        lines += ["", f"total device time in the captured region: {total:.2f} ms", ""]
    summary = {"tag": a.tag, "workload": a.workload}
    if a.rep:
        m = raw_metrics(a.rep)
        lines += [f"## `ncu --set full` of the top kernel ({pathlib.Path(a.rep).name})", "",
                  f"kernel: `{short(m.get('Kernel Name', ('?',))[0])}`", "", "| metric | value | unit |", "|---|---:|---|"]
        for k in KEEP:
            if k in m:
                lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")

        def num(key):
            v, u = m[key]
            x = float(v.replace(",", ""))
            return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
        summary["dram_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        summary["kernel_ms_under_ncu"] = float(m["gpu__time_duration.sum"][0].replace(",", ""))
        summary["smem_pipe_pct"] = float(m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"][0])
        summary["registers"] = int(float(m["launch__registers_per_thread"][0]))
    (prof / f"{a.tag}.md").write_text("\n".join(lines) + "\n")
    if a.rep:
        (prof / "scan_ncu_summary.json").write_text(json.dumps(summary, indent=1) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
