"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --no-graph` per kernel.

usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md
A "step" is delimited by the loss kernel (`head_loss_kernel`, one launch per training step); the table is the mean over the complete
steps in the capture.  ncu serialises launches and runs them cold, so use the SHARES, not the absolute times.
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"((?:clk::)?[A-Za-z_0-9:]+(?:<[^(]*>)?)", name)
    s = m.group(1) if m else name
    s = s.replace("clk::", "")
    return s[:90]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"],
                     r["Block Size"]))
    marks = [i for i, r in enumerate(rows) if r[0].startswith(("ce_kd_loss", "head_loss_kernel"))]
    if len(marks) < 3:
        raise SystemExit("need at least three steps in the capture")
    # a step spans from one loss launch to the next (forward of step k+1 precedes its loss; the sum over a
    # loss-to-loss window is exactly one backward + one optimiser + one forward)
    spans = list(zip(marks[:-1], marks[1:]))[1:]  # the first window holds the optimiser's one-time state allocation
    nstep = len(spans)
    agg = OrderedDict()
    for a, b in spans:
        for name, us, grid, block in rows[a:b]:
            e = agg.setdefault(name, [0, 0.0, grid, block])
            e[0] += 1
            e[1] += us
    total = sum(e[1] for e in agg.values())
    print(f"source: {path}; {len(rows)} launches captured, {nstep} loss-to-loss windows averaged\n")
    print(f"serialised GPU time per step: {total / nstep / 1e3:.3f} ms\n")
    print("| kernel | launches/step | us/step | share | avg us/launch | grid (last) | block |")
    print("|---|---|---|---|---|---|---|")
    for name, (n, us, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n / nstep:.1f} | {us / nstep:.1f} | {100 * us / total:.1f} % | {us / n:.1f} | {grid} | {block} |")


if __name__ == "__main__":
    main(sys.argv[1])
