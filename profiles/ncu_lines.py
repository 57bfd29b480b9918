#!/usr/bin/env python
"""Per-source-line view of an ncu report (kernel compiled with -lineinfo, captured with --import-source on).
usage: python profiles/ncu_lines.py <report.ncu-rep> [top_n]
Prints file:line, share of executed warp instructions, share of stall samples, average active threads, the top
not-issued stall reasons and the source text."""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []
fname, hdr = "?", None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = os.path.basename(r[1])
        continue
    if r and r[0] == "Line No":
        hdr = r
        ci_inst, ci_thr, ci_samp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        stall_cols = [(h.replace(" (Not Issued)", "")[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" in h]
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue                                   # SASS rows have an empty line number
    try:
        inst, thr, samp = int(r[ci_inst] or 0), int(r[ci_thr] or 0), int(r[ci_samp] or 0)
    except ValueError:
        continue
    st = []
    for name, i in stall_cols:
        try:
            v = int(r[i] or 0)
        except ValueError:
            v = 0
        if v:
            st.append((v, name))
    lines.append((inst, samp, thr, fname, int(r[0]), r[1], sorted(st, reverse=True)[:3]))
tot = sum(l[0] for l in lines) or 1
tots = sum(l[1] for l in lines) or 1
print("total warp instructions %d, stall samples %d" % (tot, tots))
for inst, samp, thr, f, ln, src, st in sorted(lines, key=lambda l: -(l[0] / tot + l[1] / tots))[:top]:
    print("%-26s %5.1f%% inst %5.1f%% samp thr %4.1f  %-34s| %s" % ("%s:%d" % (f, ln), 100 * inst / tot, 100 * samp / tots, thr / max(inst, 1),
          " ".join("%s:%d" % (n, v) for v, n in st), src.strip()[:80]))
