#!/usr/bin/env python
"""Per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics in the shipped library -> profiles/<name>.
usage: python tools/sass_summary.py [out-name]      (CPU only: cuobjdump -sass on yolo_v3_tf2_b200/lib/liby3b200.so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "yolo_v3_tf2_b200", "lib", "liby3b200.so")
name = sys.argv[1] if len(sys.argv) > 1 else "r2_sass_summary.txt"
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
rows = []
for f in funcs:
    mangled = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void ", "")
    lines = f.split("\n")
    n_ins = sum(1 for l in lines if re.search(r"/\*[0-9a-f]{4,}\*/\s+[A-Z@]", l))
    c = {p: sum(1 for l in lines if re.search(r"\b" + re.escape(p), l))
         for p in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS")}
    c["UTCHMMA.2CTA"] = sum(1 for l in lines if "UTCHMMA.2CTA" in l)
    c["IM2COL"] = sum(1 for l in lines if "IM2COL" in l)
    rows.append((dem, n_ins, c))
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
hdr = (f"{'kernel':72s} {'SASS':>6s} {'UTCHMMA':>8s} {'.2CTA':>6s} {'LDTM':>5s} {'UTMALDG':>8s} {'IM2COL':>7s} {'UTMASTG':>8s} "
       f"{'UBLKCP':>7s} {'UTCBAR':>7s} {'SYNCS':>6s}")
out = [f"# SASS summary of yolo_v3_tf2_b200/lib/liby3b200.so (release build, sm_100a), sources at commit {commit} (+ working tree)",
       "# cuobjdump -sass, per-function counts of the mnemonics B200_PROFILING.md names: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2),",
       "# LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store (IM2COL = im2col mode), UBLKCP = bulk copy,",
       "# UTCBAR = tcgen05.commit, SYNCS = mbarrier operations.  SASS = instructions in the function.", hdr]
tot = collections.Counter()
for dem, n, c in sorted(rows, key=lambda r: (-r[2]["UTCHMMA"], r[0])):
    if not any(c.values()):
        continue
    out.append(f"{dem[:72]:72s} {n:6d} {c['UTCHMMA']:8d} {c['UTCHMMA.2CTA']:6d} {c['LDTM']:5d} {c['UTMALDG']:8d} {c['IM2COL']:7d} "
               f"{c['UTMASTG']:8d} {c['UBLKCP']:7d} {c['UTCBAR']:7d} {c['SYNCS']:6d}")
    tot.update(c)
out.append(f"{'total':72s} {'':6s} {tot['UTCHMMA']:8d} {tot['UTCHMMA.2CTA']:6d} {tot['LDTM']:5d} {tot['UTMALDG']:8d} {tot['IM2COL']:7d} "
           f"{tot['UTMASTG']:8d} {tot['UBLKCP']:7d} {tot['UTCBAR']:7d} {tot['SYNCS']:6d}")
open(os.path.join(ROOT, "profiles", name), "w").write("\n".join(out) + "\n")
print("\n".join(out))
