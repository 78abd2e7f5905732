"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py gpurun_out/r1_launches.csv gpurun_out/r1_fused_full.ncu-rep r1
"""
import collections
import csv
import json
import os
import subprocess
import sys

launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")

rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}


def short(n):
    return n.split("(")[0].replace("void ", "").replace("<unnamed>::", "")


seq = [(short(r[idx["Kernel Name"]]), float(r[idx["Metric Value"]]), r[idx["Grid Size"]]) for r in rows[1:]]
# one pipeline step = points_to_cartesian, k1, fused k2, k5 (which also gathers the winner): 4 launches; the bench does 3 warm-up steps first
step_names = ["points_to_cartesian", "k1_coefficients", "k2_evaluate<1, 1,", "k5_argmin"]
n_step = len(step_names)
starts = [i for i in range(len(seq) - n_step + 1) if all(seq[i + j][0].startswith(step_names[j]) for j in range(n_step))]
lines = [f"# ncu launch list summary ({tag})", "",
         f"Source: `{os.path.basename(launches)}` (`ncu --metrics gpu__time_duration.sum --clock-control none`, same command as the",
         "bench line: `python bench.py --steps 5 --warmup 3 --no-cpu`).  Per-launch times are cold-cache and serialised:",
         "compare SHARES with the CUDA-event numbers of bench.py, not absolutes.", "",
         f"Pipeline steps found: {len(starts)} ({n_step} launches each).  Mean over steps {min(3, len(starts))}..{len(starts) - 1}:", "",
         "| kernel | grid | mean us | share of step |", "|---|---|---|---|"]
use = starts[3:] if len(starts) > 3 else starts
tot = collections.OrderedDict((n, 0.0) for n in step_names)
grid = {}
for i in use:
    for n0, (n, t, g) in zip(step_names, seq[i:i + n_step]):
        tot[n0] += t
        grid[n0] = g
step_total = sum(tot.values())
for n in step_names:
    lines.append(f"| `{n.rstrip(',')}` | {grid.get(n, '')} | {tot[n] / max(len(use), 1) / 1e3:.1f} | {100 * tot[n] / max(step_total, 1e-9):.1f} % |")
lines += ["", "All launches of the capture, by kernel:", "", "| kernel | launches | total ms | mean us |", "|---|---|---|---|"]
agg = collections.OrderedDict()
for n, t, g in seq:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"| `{n[:80]}` | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} |")
open(os.path.join(out_dir, f"{tag}_launches_summary.md"), "w").write("\n".join(lines) + "\n")

# `rep` is an .ncu-rep, or the CSV of `ncu -i <rep> --page raw --csv` when the report itself was too large to bring back
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                                                                     capture_output=True, text=True).stdout
out_name = sys.argv[4] if len(sys.argv) > 4 else f"{tag}_fused_kernel_ncu_summary.md"
note = sys.argv[5] if len(sys.argv) > 5 else "python bench.py --steps 2 --warmup 3 --no-cpu`, i.e. the bench's own 512-scenario launch"
r = list(csv.reader(raw.splitlines()))
h, units = r[0], r[1]
ix = {k: i for i, k in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size"]
md = [f"# ncu --set full summary ({tag})", "", f"Source: `{os.path.basename(rep)}` (`ncu --set full --clock-control none ...", note + ").", ""]
traffic = {}
seen = set()
for row in r[2:]:
    name = row[ix["Kernel Name"]]
    if name in seen:        # one section per distinct kernel (the first captured launch)
        continue
    seen.add(name)
    md += [f"## `{name[:100]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in ix:
            md.append(f"| {w} | {row[ix[w]]} | {units[ix[w]]} |")
    stalls = []
    for k in h:
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                stalls.append((float(row[ix[k]]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    s = sum(v for v, _ in stalls) or 1.0
    md += ["", "Warp stall samples: " + ", ".join(f"{k} {100 * v / s:.1f}%" for v, k in sorted(stalls, reverse=True)[:8]), ""]

    def to_bytes(val, unit):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        return float(val) * mult
    rd = to_bytes(row[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(row[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    key = "k2k3k4_fused" if "<1, 1," in name or "(bool)1, (bool)1" in name else ("k2k3_fused" if "<1, 0," in name else "k2_evaluate")
    traffic[key] = rd + wr
    md.append(f"DRAM traffic per launch: {rd / 1e9:.3f} GB read + {wr / 1e9:.3f} GB written = **{(rd + wr) / 1e9:.3f} GB**.")
# a hand-written "Executed SASS by region" section (from the source page of the same capture) survives a regeneration
out_path = os.path.join(out_dir, out_name)
keep = ""
if os.path.exists(out_path):
    old_text = open(out_path).read()
    at = old_text.find("## Executed SASS by region")
    if at >= 0:
        keep = "\n" + old_text[at:]
open(out_path, "w").write("\n".join(md) + "\n" + keep)
if len(sys.argv) <= 4:      # only the full-size capture feeds traffic.json (bytes per launch of the bench's own launch)
    tpath = os.path.join(out_dir, "traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old.update(traffic)
    json.dump(old, open(tpath, "w"), indent=1)
print(traffic)
