#!/usr/bin/env python
"""Kernel timeline of ONE one-shot xpress_huff_decompress call (run under `ncu --metrics gpu__time_duration.sum`).
   python profiles/oneshot_timeline.py [MiB] [ours|oracle]      # who made the stream: this library (independent chunks) or the oracle"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mscompress_b200 as ms  # noqa: E402

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 8) << 20
who = sys.argv[2] if len(sys.argv) > 2 else "ours"
plain = ms.generate("text", n, 42)
if who == "ours":
    st, comp = ms.raw_call("xpress_huff_compress", plain, n + 34 + 258 + 258 * (n // 65536))
else:
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle")], check=True)
    orc = ctypes.CDLL(os.path.join(root, "oracle", "liboracle.so"))
    out = ctypes.create_string_buffer(n + n // 8 + 4096)
    m = ctypes.c_size_t(len(out))
    assert orc.orc_xpress_huff_compress(plain, ctypes.c_size_t(n), out, ctypes.byref(m)) == 0
    comp = out.raw[:m.value]
ts = []
for k in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4):
    t0 = time.perf_counter()
    st2, back = ms.raw_call("xpress_huff_decompress", comp, n)
    ts.append(1e3 * (time.perf_counter() - t0))
    assert st2 == 0 and back == plain
print("xpress_huff_decompress of one %d MiB stream made by %s: %s ms -> best %.2f ms = %.1f MB/s" % (n >> 20, who, " ".join("%.1f" % t for t in ts), min(ts), n / min(ts) / 1e3))
