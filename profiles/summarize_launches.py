"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_*.md"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)   # -> microseconds
        m = re.search(r"([A-Za-z_][A-Za-z0-9_]*_kernel|[A-Za-z_][A-Za-z0-9_]*)\s*(<[^(]*>)?\s*\(", r["Kernel Name"].replace("<unnamed>::", ""))
        name = m.group(1) if m else r["Kernel Name"][:60]
        rows.append((name, v))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    print(f"# launch list summary: {path}\n")
    print(f"{len(rows)} launches, {tot/1e3:.2f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {v/1e3:.3f} | {100*v/tot:.1f}% | {v/c:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
