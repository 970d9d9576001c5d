"""Summarise an `ncu --set full` raw-page CSV (`ncu -i X.ncu-rep --page raw --csv`) launch by launch.

usage: python scripts/summarize_full.py gpurun_out/pair_full_raw.csv [--json profiles/<name>.json] > profiles/<name>.md

--json writes {"mean_dram_bytes_per_launch": ..., "launches": ..., "kernels": {name: {...}}} over the launches whose kernel
name matches --match (default: every launch); bench.py reads `roofline.traffic` from that file.
"""
import csv
import re
import sys

COLS = [
    ("us", "gpu__time_duration.sum", "time"),
    ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    ("smem-port %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("DRAM rd MB", "dram__bytes_read.sum", None),
    ("DRAM wr MB", "dram__bytes_write.sum", None),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(path, json_path=None, match=None):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"source: {path}, {len(data)} launches\n")
    print("| # | kernel | grid | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    tot_bytes, n = 0.0, 0
    sel_bytes, sel_n, per = 0.0, 0, {}
    for k, r in enumerate(data):
        name = re.sub(r"^void (clk::)?", "", r[ix["Kernel Name"]]).split("(")[0]
        row_bytes, row_us = 0.0, 0.0
        cells = []
        for label, key, scale in COLS:
            v = r[ix[key]]
            if scale == "time":
                t = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[units[ix[key]]]
                row_us = t
                cells.append(f"{t:.1f}")
            elif scale is None:
                b = to_bytes(v, units[ix[key]])
                tot_bytes += b
                row_bytes += b
                cells.append(f"{b / 1e6:.1f}")
            else:
                cells.append(f"{float(v.replace(',', '')) * scale:.1f}" if v else "-")
        n += 1
        if match is None or re.search(match, name):
            sel_bytes += row_bytes
            sel_n += 1
            e = per.setdefault(name, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
            e["launches"] += 1
            e["us"] += row_us
            e["dram_bytes"] += row_bytes
        print(f"| {k} | `{name}` | {r[ix['Grid Size']]} | " + " | ".join(cells) + " |")
    print(f"\nmean DRAM traffic per launch (read + write): {tot_bytes / n / 1e6:.1f} MB")
    if match is not None and sel_n:
        print(f"mean DRAM traffic per launch of the kernels matching /{match}/ ({sel_n} launches): {sel_bytes / sel_n / 1e6:.1f} MB")
    if json_path:
        import json
        json.dump({"source": path, "match": match, "launches": sel_n, "mean_dram_bytes_per_launch": sel_bytes / max(sel_n, 1),
                   "kernels": per}, open(json_path, "w"), indent=1)


if __name__ == "__main__":
    a = sys.argv[1:]
    jp = a[a.index("--json") + 1] if "--json" in a else None
    mt = a[a.index("--match") + 1] if "--match" in a else None
    main(a[0], jp, mt)
