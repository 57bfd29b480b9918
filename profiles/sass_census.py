#!/usr/bin/env python
"""SASS mnemonic census of the product library: python profiles/sass_census.py > profiles/<round>_sass_census.txt

Counts, per kernel of ms-compress_b200/lib/libMSCompression.so (cuobjdump -sass, sm_100a), the instructions that say how data
moves: bulk async copies (UBLKCP) and mbarrier operations (SYNCS), 16-/8-byte and byte-wide global accesses, 16-byte
shared-memory accesses, shared atomics, warp match / reduce.  Runs without a GPU.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ms-compress_b200", "lib", "libMSCompression.so")
WATCH = ["UBLKCP", "SYNCS", "LDG.E.U8", "STG.E.U8", "LDG.E.64", "STG.E.64", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128",
         "ATOMS", "MATCH.ANY", "REDUX"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                # LDG.E.U8.CONSTANT etc. count as LDG.E.U8; LDG.E.64 must not swallow LDG.E.U8
                per[cur][w] += 1
                break
    names = list(per)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangle[n] = re.sub(r"\(.*", "", re.sub(r"^void ", "", d))
    merged = collections.OrderedDict()
    for n in names:                        # template instances of one kernel are added up
        k = re.sub(r"<.*", "", demangle[n])
        merged.setdefault(k, collections.Counter()).update(per[n])
    total = collections.Counter()
    for c in merged.values():
        total.update(c)
    print("# SASS mnemonic census of ms-compress_b200/lib/libMSCompression.so (cuobjdump -sass, sm_100a); profiles/sass_census.py")
    print("# UBLKCP = cp.async.bulk (bulk global<->shared copies), SYNCS = mbarrier operations; template instances of a kernel are summed")
    print("total: " + " ".join("%s=%d" % (w, total[w]) for w in WATCH if total[w]))
    for k, c in sorted(merged.items(), key=lambda kv: -(kv[1]["LDG.E.U8"] + kv[1]["STG.E.U8"])):
        if not sum(c.values()):
            continue
        print("%-44s %s" % (k, " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])))


if __name__ == "__main__":
    sys.exit(main())
