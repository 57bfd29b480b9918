#!/usr/bin/env python
"""One-shot C-ABI calls on ONE buffer (host memory): the case a drop-in user of the reference hits first.
   python profiles/oneshot_probe.py [MiB]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mscompress_b200 as ms  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = mib << 20
plain = ms.generate("text", n, 42)
for fmt, maxsz in (("lznt1", n + 3 + 2 * ((n + 4095) // 4096)), ("xpress", n + 4 + 4 * (n // 32)), ("xpress_huff", n + 34 + 258 + 258 * (n // 65536))):
    tc, td = [], []
    for k in range(3):
        t0 = time.perf_counter()
        st, comp = ms.raw_call(fmt + "_compress", plain, maxsz)
        t1 = time.perf_counter()
        st2, back = ms.raw_call(fmt + "_decompress", comp, n)
        t2 = time.perf_counter()
        assert st == 0 and st2 == 0 and back == plain
        tc.append(t1 - t0); td.append(t2 - t1)
    print(f"{fmt:12s} {mib} MiB one buffer: compress {1e3 * min(tc):8.2f} ms ({n / min(tc) / 1e6:8.1f} MB/s)   decompress {1e3 * min(td):8.2f} ms ({n / min(td) / 1e6:8.1f} MB/s)")
