#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction metrics by CUDA source line.
usage: python profiles/ncu_lines.py <report.ncu-rep> [top_n]   (needs the kernel compiled with -lineinfo and
captured with --import-source on).  Prints: line, instructions executed, share, avg active threads, top stall
reasons (sampled, not-issued), source text."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find header row
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
inst_c, thr_c, samp_c = ci["Instructions Executed"], ci["Thread Instructions Executed"], ci["# Samples"]
stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" in h]
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int), ""])
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[ci["Source"]]
    addr = r[ci["Address"]] if "Address" in ci else ""
    # cuda,sass view: CUDA source rows carry the line text; SASS rows follow them
    key = r[0]
    try:
        n = int(r[inst_c] or 0)
    except ValueError:
        continue
    if not addr.startswith("0x") and not key.startswith("0x"):
        cur = (key, src)
        continue
    if cur is None:
        continue
    a = agg[cur]
    a[0] += n
    a[1] += int(r[thr_c] or 0)
    a[2] += int(r[samp_c] or 0)
    for h, i in stall_cols:
        try:
            a[3][h.replace(" (Not Issued)", "")] += int(r[i] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
for (line, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print("%6s %6.2f%% inst %6.2f%% samp  thr %5.1f  %-44s | %s" % (line, 100 * a[0] / tot, 100 * a[2] / tots, a[1] / max(a[0], 1),
          " ".join("%s:%d" % (k[6:], v) for k, v in st if v), src.strip()[:90]))
