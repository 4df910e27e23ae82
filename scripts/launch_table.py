#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def table(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] == "gpu__time_duration.sum":
                v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
                v = v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)
                name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
                agg[name + " grid=" + d.get("Grid Size", "")].append(v)
    return agg


if __name__ == "__main__":
    agg = table(sys.argv[1])
    tot = sum(sum(v) for v in agg.values())
    print(f"| kernel | launches | total ms | each us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v) / 1e3:.3f} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.2f}% |")
