#!/usr/bin/env python
"""Turns gpurun_out/<tag>_launches.csv (+ optional .ncu-rep files) into the tracked summaries under profiles/.
usage: python tools_summarize_profiles.py <tag> [<out-name>]"""
import collections
import csv
import io
import os
import subprocess
import sys

tag = sys.argv[1]
name = sys.argv[2] if len(sys.argv) > 2 else tag
os.makedirs("profiles", exist_ok=True)
out = []
lp = f"gpurun_out/{tag}_launches.csv"
if os.path.exists(lp):
    txt = open(lp).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID","Process ID"'):])))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = r["Kernel Name"].split("(")[0].replace("void ", "")
        tot[n][0] += 1
        tot[n][1] += float(r["Metric Value"]) / 1e3
    T = sum(v[1] for v in tot.values())
    out.append(f"## ncu launch list ({len(rows)} launches, `--metrics gpu__time_duration.sum --clock-control none`)\n")
    out.append("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
    out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{n}` | {c} | {t:.1f} | {100 * t / T:.1f}% |")
    out.append(f"| total | {len(rows)} | {T:.1f} | 100% |\n")
    with open(f"profiles/{name}_launches.csv", "w") as f:
        f.write("kernel,grid,block,duration_ns\n")
        for r in rows:
            f.write(f"\"{r['Kernel Name'].split('(')[0]}\",\"{r['Grid Size']}\",\"{r['Block Size']}\",{r['Metric Value']}\n")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
for rep in sorted(f for f in os.listdir("gpurun_out") if f.startswith(tag + "_prof") and f.endswith(".ncu-rep")):
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    out.append(f"## `ncu --set full` capture {rep}\n")
    for r in rows[2:]:
        kn = r[hdr.index("Kernel Name")]
        out.append(f"### {kn}  grid {r[hdr.index('launch__grid_size')]} x block {r[hdr.index('launch__block_size')]}\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for w in want:
            if w in hdr:
                out.append(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
        for i, h in enumerate(hdr):
            if "pipe_tensor" in h and h not in want:
                out.append(f"| {h} | {r[i]} | {units[i]} |")
        out.append("")
open(f"profiles/{name}.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
