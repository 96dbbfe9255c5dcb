#!/usr/bin/env python
"""Reads an .ncu-rep (captured with --set full --import-source on) and prints, per kernel: headline metrics and the
stall-sample breakdown of the warp roles, identified by the barrier/TMA/MMA instructions around them."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__cycles_elapsed.avg.per_second']
names = []
for r in rows[2:]:
    names.append(r[hdr.index('Kernel Name')])
    print('=====', r[hdr.index('Kernel Name')][:70], 'grid', r[hdr.index('launch__grid_size')])
    for k in keys:
        if k in hdr:
            print(f"  {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
for bi, b in enumerate(blocks):
    h = b['rows'][0]
    data = [r for r in b['rows'][1:] if len(r) > 5]
    iS, iSrc, iEx = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    tot = sum(int(r[iS] or 0) for r in data)
    print(f"\n----- [{bi}] {b['name'][:60]}  total samples {tot}")
    for i, r in enumerate(data):
        s = int(r[iS] or 0)
        if s > tot * 0.01 or any(k in r[iSrc] for k in ('UTCHMMA', 'UTMALDG', 'UTCBAR', 'LDTM')):
            print(f"{i:5d} smp={s:6d} ({100*s/tot:4.1f}%) exec={r[iEx]:>8s} {r[iSrc][:95]}")
