#!/usr/bin/env python
"""Shared-memory wavefronts (actual vs ideal) of a kernel in an ncu report, by CUDA source line.
usage: tools/ncu_smem_by_line.py <report.ncu-rep> <kernel-name-substring> [lib.so] [top] [cubin-symbol-substring]"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, kname = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "vorbis_b200", "libvorbis_b200.so")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
cname = sys.argv[5] if len(sys.argv) > 5 else kname   # mangled-name substring in the cubin (template instances)
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
lines = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and cname in l][0]
off2line, cur = {}, ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname],
                     capture_output=True, text=True).stdout
hdr, data, seen = None, [], set()
for r in csv.reader(out.split("\n")):
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in seen:
        seen.add(r[0]); data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ci["Address"]], 16)
agg = defaultdict(lambda: [0, 0, 0, ""])
for r in data:
    off = int(r[ci["Address"]], 16) - base
    fl, sass = off2line.get(off, (("?", 0), ""))
    w = int(r[ci["L1 Wavefronts Shared"]] or 0); wi = int(r[ci["L1 Wavefronts Shared Ideal"]] or 0)
    if not w:
        continue
    a = agg[(fl, off)]
    a[0] += w; a[1] += wi; a[2] += int(r[ci["Instructions Executed"]] or 0); a[3] = sass[:60]
tw = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print("shared wavefronts %d, ideal %d" % (tw, ti))
for (fl, off), a in sorted(agg.items(), key=lambda kv: -(kv[1][0] - kv[1][1]))[:top]:
    print("%-24s off %05x  wavefronts %9d ideal %9d  x%.1f  inst %8d  %s" % ("%s:%d" % fl, off, a[0], a[1], a[0] / max(a[1], 1), a[2], a[3]))
