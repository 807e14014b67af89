#!/usr/bin/env python
"""List the hot SASS regions of a kernel in an ncu report with the CUDA source lines they come from.
usage: tools/ncu_regions.py <report.ncu-rep> <kernel-substring-in-cubin-symbol> <matching lib.so> [rows] [min_exec] [cubin-symbol-substring]
The .so MUST be the build the report was captured from."""
import csv, os, re, subprocess, sys, tempfile
rep, kname, so = sys.argv[1:4]
rows_n = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 2e5
cname = sys.argv[6] if len(sys.argv) > 6 else kname   # mangled-name substring in the cubin (template instances)
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and cname in l][0]
off2 = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        off2[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname], capture_output=True, text=True).stdout
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
recs = [(int(r[ci["Address"]], 16) - base, int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)) for r in data]
tot = sum(r[1] for r in recs)
print("total executed %.1fM  (%.0f per row)" % (tot / 1e6, tot / rows_n))
i = 0
while i < len(recs):
    if recs[i][1] >= thr:
        j = i
        while j < len(recs) and recs[j][1] >= thr:
            j += 1
        s = sum(r[1] for r in recs[i:j])
        smp = sum(r[2] for r in recs[i:j])
        if s > tot * 0.01:
            locs = {}
            for r in recs[i:j]:
                k = off2.get(r[0])
                locs[k] = locs.get(k, 0) + r[1]
            top = sorted(locs.items(), key=lambda kv: -kv[1])[:3]
            print("off %05x-%05x n=%3d  %5.1f%% of instr (%6.0f/row) samples %5d  lines %s"
                  % (recs[i][0], recs[j - 1][0], j - i, 100.0 * s / tot, s / rows_n, smp,
                     ", ".join("%s:%d" % k for k, _ in top if k)))
        i = j
    else:
        i += 1
