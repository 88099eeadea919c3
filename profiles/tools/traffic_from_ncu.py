"""Turn an ncu CSV of profiles/tools/plan_once.py (metrics gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum) into (a) a per-kernel table (markdown on stdout) and (b) an entry of profiles/r02_traffic.json,
which bench.py reads for `roofline.traffic`.

    python profiles/tools/traffic_from_ncu.py gpurun_out/plan.csv ELIC_united 480 640 8 bf16 [algorithmic_conv_bytes]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse(path):
    rows = {}
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        k = rows.setdefault(r["ID"], {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        name = r["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)       # -> us
        k[name] = v
    return list(rows.values())


def main():
    path, model, H, W, B, precision = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    alg = int(sys.argv[7]) if len(sys.argv) > 7 else None
    rows = parse(path)
    agg = {}
    for r in rows:
        base = r["name"].replace("void ", "").replace("<unnamed>::", "").split("(")[0].split("<")[0]
        a = agg.setdefault(base, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += r.get("gpu__time_duration.sum", 0.0)
        a[2] += r.get("dram__bytes_read.sum", 0.0)
        a[3] += r.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | share | dram read MB | dram write MB |\n|---|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} |")
    conv = [a for k, a in agg.items() if k.startswith("conv_") or k.startswith("rb_fused")]
    dram = sum(a[2] + a[3] for a in conv)
    out = os.path.join(ROOT, "profiles", "r02_traffic.json")
    entries = json.load(open(out)) if os.path.exists(out) else []
    entries = [e for e in entries if (e["model"], e["height"], e["width"], e["precision"]) != (model, H, W, precision)]
    entries.append({"model": model, "height": H, "width": W, "precision": precision, "batch": B,
                    "conv_launches": sum(a[0] for a in conv), "dram_bytes_per_pair": dram / B,
                    "algorithmic_bytes_per_pair": alg / B if alg else None,
                    "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the conv launches of one encoder + decoder plan "
                              f"(profiles/tools/plan_once.py {model} {H} {W} {B} {precision}), cold cache, serialised"})
    json.dump(entries, open(out, "w"), indent=1)
    print(f"\nconv DRAM traffic: {dram / B / 1e9:.3f} GB per pair" + (f" (algorithmic {alg / B / 1e9:.3f} GB)" if alg else ""))


if __name__ == "__main__":
    main()
