#!/usr/bin/env python
"""Per-kernel time of the GPU chunk index on the compressed form of a text stream.  python profiles/index_probe.py [MiB]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mscompress_b200 as ms  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = ms.GpuContext()
host = np.empty(mib << 20, dtype=np.uint8)
ms.generate_batch_into(host, 1 << 20, mib, 42, "text")
d_in = torch.from_numpy(host).to(ctx.device)
d_comp, d_off, d_res = ctx.lznt1_compress(d_in)
C = ms.parse_result(d_res)["bytes"]
nch = (mib << 20) // 4096
for rep in range(2):
    ctx.profile(True)
    d_idx, info = ctx.lznt1_index(d_comp, chunk_cap=nch + 8, in_len=C)
    rep_txt = ctx.profile_report()
assert info["nchunks"] == nch and info["tail"] == 0 and info["consumed"] == C, info
assert torch.equal(d_idx[:nch], d_off[:nch].to(torch.int64))
print(f"{C} compressed bytes, {nch} chunks")
print(rep_txt)
