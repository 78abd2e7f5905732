"""Per-source-line summary of an ncu source page exported with
    ncu -i REPORT --page source --csv --print-source sass,cuda > page.csv
(instructions executed and warp-stall samples per CUDA source line, grouped into line ranges given on the command line).

    python scripts/ncu_lines.py page.csv [--top 40] [--ranges name:lo-hi,name:lo-hi,...]
"""
import argparse
import csv
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--ranges", default="")
a = ap.parse_args()
rows = list(csv.reader(open(a.csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
ci, cs = H.index("Instructions Executed"), H.index("# Samples")
lines = {}
for r in rows[hdr + 1:]:
    if r and r[0].isdigit() and r[2] == "-":
        ln = int(r[0])
        ins = int(r[ci]) if r[ci].isdigit() else 0
        smp = int(r[cs]) if r[cs].isdigit() else 0
        e = lines.setdefault(ln, [0, 0, r[1]])
        e[0] += ins
        e[1] += smp
ti, ts = sum(v[0] for v in lines.values()), sum(v[1] for v in lines.values())
print(f"total warp instructions {ti:,}  samples {ts:,}")
if a.ranges:
    for spec in a.ranges.split(","):
        name, rng = spec.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        i = sum(v[0] for k, v in lines.items() if lo <= k <= hi)
        s = sum(v[1] for k, v in lines.items() if lo <= k <= hi)
        print(f"{name:28s} lines {lo:5d}-{hi:5d}  instr {i:14,} {100 * i / ti:5.1f} %   samples {s:8,} {100 * s / max(ts, 1):5.1f} %")
print("top lines by instructions:")
for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:a.top]:
    print(f"{ln:5d} {v[0]:13,} {100 * v[0] / ti:5.1f}% smp {100 * v[1] / max(ts, 1):5.1f}%  {v[2][:110]}")
