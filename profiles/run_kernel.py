#!/usr/bin/env python
"""Small driver used under ncu: runs encode+decode of one format a few times on synthetic text.
    python profiles/run_kernel.py lznt1 128            # 128 MiB stream
    python profiles/run_kernel.py xpress_huff 4096     # 4096 buffers x 64 KiB
    python profiles/run_kernel.py xpress 256           # 256 buffers x 1 MiB
Prints per-kernel CUDA-event times (not valid under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mscompress_b200 as ms  # noqa: E402

fmt, count = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kind = sys.argv[4] if len(sys.argv) > 4 else "text"
ctx = ms.GpuContext()
dev = ctx.device
if fmt == "lznt1":
    n = count << 20
    host = np.empty(n, dtype=np.uint8)
    ms.generate_batch_into(host, 1 << 20, count, 42, kind)
    d_in = torch.from_numpy(host).to(dev)
    ctx.profile(True)
    for _ in range(reps):
        d_comp, d_off, d_res = ctx.lznt1_compress(d_in)
        d_out, d_res2 = ctx.lznt1_decompress(d_comp, d_off, n // 4096)
    torch.cuda.synchronize()
    assert torch.equal(d_out[:n], d_in)
    U, C = n, ms.parse_result(d_res)["bytes"]
else:
    each = 65536 if fmt == "xpress_huff" else 1 << 20
    capf = {"xpress_huff": lambda n: n + 34 + 258 + 258 * (n // 65536), "xpress": lambda n: n + 4 + 4 * (n // 32)}[fmt]
    U = each * count
    host = np.empty(U, dtype=np.uint8)
    ms.generate_batch_into(host, each, count, 42, kind)
    d_in = torch.from_numpy(host).to(dev)
    slot = (capf(each) + 15) // 16 * 16
    d_comp = torch.empty(slot * count, dtype=torch.uint8, device=dev)
    d_out = torch.empty(U, dtype=torch.uint8, device=dev)
    ar = torch.arange(count, dtype=torch.int64, device=dev)
    in_off, c_off = ar * each, ar * slot
    in_len = torch.full((count,), each, dtype=torch.int32, device=dev)
    c_cap = torch.full((count,), slot, dtype=torch.int32, device=dev)
    c_len = torch.zeros(count, dtype=torch.int32, device=dev)
    st1 = torch.zeros(count, dtype=torch.int32, device=dev)
    o_len = torch.zeros(count, dtype=torch.int32, device=dev)
    st2 = torch.zeros(count, dtype=torch.int32, device=dev)
    enc = ms.Batch(count, in_off.data_ptr(), in_len.data_ptr(), c_off.data_ptr(), c_cap.data_ptr(), c_len.data_ptr(), st1.data_ptr())
    dec = ms.Batch(count, c_off.data_ptr(), c_len.data_ptr(), in_off.data_ptr(), in_len.data_ptr(), o_len.data_ptr(), st2.data_ptr())
    ctx.profile(True)
    indexed = fmt == "xpress"                  # plain Xpress: with the side-band restart index (one entry per 16 KiB)
    seg_bytes, stride = 16384, (each + 16383) // 16384
    d_index = ctx.xpress_index_tensor(count, stride) if indexed else None
    xh = ctx.xhuff_index_alloc([each] * count) if fmt == "xpress_huff" else None
    for _ in range(reps):
        if indexed:
            ctx.xpress_compress_indexed(d_in, d_comp, enc, seg_bytes, d_index, stride)
            ctx.xpress_decompress_indexed(d_comp, d_out, dec, d_index, stride)
        else:                                      # Xpress-Huffman: the encoder writes the restart index, the decoder starts from it
            ctx.xhuff_compress_indexed(d_in, d_comp, enc, *xh)
            d_out.zero_()
            ctx.xhuff_decompress_indexed(d_comp, d_out, dec, *xh)
            if os.environ.get("RUN_UNINDEXED"):    # the index-free path (what a reference-made stream takes)
                ctx.run_batch(fmt, "decompress", d_comp, d_out, dec)
    torch.cuda.synchronize()
    assert int(st1.abs().sum()) == 0 and int(st2.abs().sum()) == 0 and torch.equal(d_out, d_in)
    C = int(c_len.to(torch.int64).sum())
if fmt == "xpress_huff":
    print("second chances:", ctx.xhuff_second_chances())
for name, (calls, ms_tot) in ctx.profile_report().items():
    per = ms_tot / calls
    print("%-24s %8.3f ms/launch  %8.1f GB/s (U+C)  %8.1f GB/s (U)" % (name, per, (U + C) / 1e9 / (per / 1e3), U / 1e9 / (per / 1e3)))
print("U=%d C=%d ratio=%.3f" % (U, C, C / U))
