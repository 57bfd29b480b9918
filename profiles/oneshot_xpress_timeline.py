#!/usr/bin/env python
"""Kernel timeline of ONE one-shot xpress_decompress call (run under `ncu --metrics gpu__time_duration.sum`).
   python profiles/oneshot_xpress_timeline.py [MiB]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mscompress_b200 as ms  # noqa: E402

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 8) << 20
plain = ms.generate("text", n, 42)
st, comp = ms.raw_call("xpress_compress", plain, n + 4 + 4 * (n // 32))
ts = []
for k in range(3):
    t0 = time.perf_counter()
    st2, back = ms.raw_call("xpress_decompress", comp, n)
    ts.append(1e3 * (time.perf_counter() - t0))
    assert st2 == 0 and back == plain
print("xpress_decompress of one %d MiB stream: %s ms -> best %.2f ms = %.1f MB/s" % (n >> 20, " ".join("%.1f" % t for t in ts), min(ts), n / min(ts) / 1e3))
