#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

usage: tools/ncu_by_line.py <report.ncu-rep> <kernel-name-substring> [lib.so] [top] [cubin-symbol-substring]
Needs -lineinfo at compile time.  Joins `ncu --page source --csv` (SASS rows) with the
line table printed by `nvdisasm -g` for the same cubin (the built .so must match the report).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, kname = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "vorbis_b200", "libvorbis_b200.so")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cname = sys.argv[5] if len(sys.argv) > 5 else kname   # mangled-name substring in the cubin (template instances)

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = dis.split("\n")
start = None
for i, l in enumerate(lines):
    if l.startswith(".text.") and cname in l:
        start = i
        break
assert start is not None, "kernel not found in cubin"
off2line = {}
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = None
data = []
seen = set()
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] not in seen:
        seen.add(r[0])
        data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ci["Address"]], 16)
agg = defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r in data:
    off = int(r[ci["Address"]], 16) - base
    (fl, _sass) = off2line.get(off, (("?", 0), ""))
    ie = int(r[ci["Instructions Executed"]] or 0)
    te = int(r[ci["Thread Instructions Executed"]] or 0)
    sm = int(r[ci["# Samples"]] or 0)
    a = agg[fl]
    a[0] += ie; a[1] += te; a[2] += sm
    tot_i += ie; tot_s += sm
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
print("%-28s %12s %7s %9s %7s %6s" % ("file:line", "warp-inst", "%inst", "samples", "%smpl", "thr/w"))
for fl, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-28s %12d %6.2f%% %9d %6.2f%% %6.1f" % ("%s:%d" % fl, a[0], 100.0 * a[0] / max(tot_i, 1), a[2],
                                                   100.0 * a[2] / max(tot_s, 1), a[1] / max(a[0], 1)))
