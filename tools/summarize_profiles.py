#!/usr/bin/env python
"""Turns gpurun_out/<tag>_launches.csv (+ optional .ncu-rep files) into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <tag> [<out-name>]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
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

tp = f"gpurun_out/{tag}_traffic.csv"
if os.path.exists(tp):
    txt = open(tp).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID","Process ID"'):])))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "")})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if "byte" in u:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        elif u in ("us", "usecond"):
            v *= 1e3
        elif u in ("ms", "msecond"):
            v *= 1e6
        d[r["Metric Name"]] = v
    # the capture holds every conv launch of several forward passes: keep the LAST pass (it starts at the stem kernel)
    items = list(per.values())
    stems = [i for i, d in enumerate(items) if "conv_stem" in d["name"] or "conv_first" in d["name"] or
             ("conv_gather_kernel<32, 128" in d["name"])]
    if stems:
        items = items[stems[-1]:]
    per = collections.OrderedDict((i, d) for i, d in enumerate(items))
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
    tt = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
    out.append(f"## DRAM traffic of one forward pass ({len(per)} conv launches -- the persistent multi-layer launches (conv_chain_kernel) cover 50 / 6 / 6 layers each --, batch 64, 416x416, uint8 input)\n")
    out.append(f"read {rd / 1e9:.3f} GB + write {wr / 1e9:.3f} GB = **{(rd + wr) / 1e9:.3f} GB** per forward pass "
               f"(algorithmic bytes, SURVEY 8d: 190 MB/img x 64 = 12.16 GB); serialised kernel time {tt / 1e6:.3f} ms\n")
    out.append("| # | kernel | us | DRAM read MB | DRAM write MB | tensor pipe % of elapsed |\n|---|---|---|---|---|---|")
    for i, d in enumerate(per.values()):
        out.append(f"| {i} | `{d['name']}` | {d.get('gpu__time_duration.sum', 0) / 1e3:.1f} | {d.get('dram__bytes_read.sum', 0) / 1e6:.1f} | "
                   f"{d.get('dram__bytes_write.sum', 0) / 1e6:.1f} | {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):.1f} |")
    out.append("")
    import json
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {"dram_bytes_read": rd, "dram_bytes_write": wr, "launches": len(per), "kernel_time_ms": tt / 1e6,
           "size": 416, "batch": 64, "commit": commit, "capture": f"gpurun_out/{tag}_traffic.csv",
           "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... -k regex:conv_ "
                      "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu (last forward pass of the capture)"}
    json.dump(rec, open(f"profiles/{name}_traffic.json", "w"), indent=1)

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
