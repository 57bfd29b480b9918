#!/usr/bin/env python
"""Latency of the one-shot C-ABI calls on small host buffers (median of 200 calls)."""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes  # noqa: E402
import mscompress_b200 as ms  # noqa: E402

L = ms.lib()
for n in (4096, 65536, 1 << 20):
    plain = ms.generate("text", n, 1)
    for fmt in ("lznt1", "xpress", "xpress_huff"):
        cap = n + n // 8 + 1024
        comp = ctypes.create_string_buffer(cap)
        back = ctypes.create_string_buffer(n)
        tc, td = [], []
        for k in range(220):
            cl = ctypes.c_size_t(cap)
            t0 = time.perf_counter()
            st = getattr(L, fmt + "_compress")(plain, ctypes.c_size_t(n), comp, ctypes.byref(cl))
            t1 = time.perf_counter()
            bl = ctypes.c_size_t(n)
            st2 = getattr(L, fmt + "_decompress")(comp, cl, back, ctypes.byref(bl))
            t2 = time.perf_counter()
            assert st == 0 and st2 == 0 and bl.value == n
            if k >= 20:
                tc.append(t1 - t0); td.append(t2 - t1)
        print(f"{fmt:12s} {n:8d} B: compress {1e6 * statistics.median(tc):8.1f} us   decompress {1e6 * statistics.median(td):8.1f} us")
