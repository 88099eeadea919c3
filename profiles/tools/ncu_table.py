"""Markdown table of the metrics that matter from an `ncu --set full` report:

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x_raw.csv ; python profiles/tools/ncu_table.py x_raw.csv "label 1" "label 2" ...
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "duration us", 1.0), ("dram__bytes_read.sum", "dram read MB", None), ("dram__bytes_write.sum", "dram write MB", None),
        ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %", 1.0),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts M", 1e-6),
        ("launch__grid_size", "grid", 1.0), ("launch__registers_per_thread", "regs", 1.0)]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    labels = sys.argv[2:]
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for k, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        cells = []
        for key, _, scale in COLS:
            if key not in idx:
                cells.append("-")
                continue
            if not r[idx[key]].strip():
                cells.append("-")
                continue
            v = float(r[idx[key]].replace(",", ""))
            u = units[idx[key]]
            if scale is None:       # bytes -> MB
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            else:
                if key == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                v *= scale
            cells.append(f"{v:.1f}")
        label = labels[k] if k < len(labels) else name
        print(f"| {label} (`{name}`) | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
