#!/usr/bin/env python
"""Times lznt1_compress and lznt1_decompress (C ABI, pinned host buffers) separately on a 1 GiB text stream.
   python profiles/e2e_probe.py [MiB]      (the MSCOMP_PIPE_* environment selects the host-wrapper configuration)"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mscompress_b200 as ms  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
U = mib << 20
host = torch.empty(U, dtype=torch.uint8).pin_memory()
ms.generate_batch_into(host.numpy(), 1 << 20, mib, 42, "text")
h_comp = torch.empty(U + U // 2048 + 64, dtype=torch.uint8).pin_memory()
h_out = torch.empty(U, dtype=torch.uint8).pin_memory()
L = ms.lib()
tc, td = [], []
for k in range(4):
    n = ctypes.c_size_t(h_comp.numel())
    t0 = time.perf_counter()
    st = L.lznt1_compress(host.data_ptr(), U, h_comp.data_ptr(), ctypes.byref(n))
    t1 = time.perf_counter()
    m = ctypes.c_size_t(U)
    st2 = L.lznt1_decompress(h_comp.data_ptr(), n.value, h_out.data_ptr(), ctypes.byref(m))
    t2 = time.perf_counter()
    assert st == 0 and st2 == 0 and m.value == U
    if k:
        tc.append(t1 - t0); td.append(t2 - t1)
assert torch.equal(h_out, host)
cfg = {k: v for k, v in os.environ.items() if k.startswith("MSCOMP_")}
print(f"{mib} MiB {cfg}: compress {1e3 * min(tc):.1f} ms, decompress {1e3 * min(td):.1f} ms (C = {n.value / U:.3f} U)")
