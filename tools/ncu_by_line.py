#!/usr/bin/env python
"""Per-source-line dynamic instruction counts and stall samples of one kernel: joins the SASS page of an ncu report with
the line table of the object file (nvdisasm -g).   python tools/ncu_by_line.py <rep> <object.o> <mangled substring> <updates>"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, sub, updates = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], float(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(tmp + "/*.cubin")[0]], capture_output=True, text=True).stdout.split("\n")
# line table of the wanted function: one source line per instruction, in address order
lines, cur, infn = [], ('?', 0), False
for l in dis:
    if l.startswith(".text."):
        infn = sub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iE, iP = hdr.index("Instructions Executed"), hdr.index("# Samples")
cnt, smp = collections.Counter(), collections.Counter()
k = 0
for r in rows[2:]:
    try:
        e, s = int(r[iE]), int(r[iP])
    except (ValueError, IndexError):
        continue
    ln = lines[k] if k < len(lines) else ('?', -1)
    k += 1
    cnt[ln] += e
    smp[ln] += s
print("instructions matched: %d of %d" % (min(k, len(lines)), k))
csrc = os.path.join(os.path.dirname(obj), "..", "csrc")
texts = {}
tot_s = sum(smp.values())
for ln in sorted(cnt):
    if cnt[ln] / updates >= float(os.environ.get("MINCNT", "0.15")) or smp[ln] / tot_s > 0.01:
        f, n = ln
        if f not in texts:
            try:
                texts[f] = open(os.path.join(csrc, f)).read().split("\n")
            except OSError:
                texts[f] = []
        t = texts[f][n - 1].strip()[:95] if 0 < n <= len(texts[f]) else "?"
        print("%7.2f /upd %5.1f%% smp  %s:%-4d %s" % (cnt[ln] / updates, 100.0 * smp[ln] / tot_s, f[:12], n, t))
