"""BASELINE.json-sized workloads through size-independent properties (the oracle cannot run 1-4 GiB in seconds):
encode -> decode round trip is the identity, per-block checksums of the decoded data equal those of the input,
sampled units decode with the oracle, and LZNT1's compressed bytes equal the oracle's on sampled 1 MiB blocks."""
import hashlib
import random

import pytest
from conftest import MAXSZ

pytestmark = pytest.mark.gpu


def _checksum_of_checksums(t, block):
    import torch
    v = t.view(-1, block).to(torch.int64)
    w = torch.arange(1, block + 1, dtype=torch.int64, device=t.device)
    return int(((v * w).sum(dim=1) % 1000000007).sum().item())


def test_lznt1_1gib_stream(ms, oracle):
    import numpy as np
    import torch
    n = 1 << 30
    host = np.empty(n, dtype=np.uint8)
    ms.generate_batch_into(host, 1 << 20, n >> 20, 42, "text")
    ctx = ms.GpuContext()
    d_in = torch.from_numpy(host).cuda()
    d_comp, d_off, d_res = ctx.lznt1_compress(d_in)
    res = ms.parse_result(d_res)
    assert res["status"] == 0
    nchunks = n // 4096
    offs = d_off.cpu().numpy()
    d_out, d_res2 = ctx.lznt1_decompress(d_comp, d_off, nchunks, in_len=res["bytes"])
    assert ms.parse_result(d_res2) == {"bytes": n, "status": 0, "first_bad": 0, "short_units": 0}
    assert torch.equal(d_out[:n], d_in)
    assert _checksum_of_checksums(d_out[:n], 4096) == _checksum_of_checksums(d_in, 4096)
    comp = d_comp[:res["bytes"]].cpu().numpy()
    rnd = random.Random(1)
    for blk in rnd.sample(range(1024), 6):          # the reference would produce exactly these bytes
        lo, hi = int(offs[blk * 256]), int(offs[(blk + 1) * 256])
        assert comp[lo:hi].tobytes() == oracle.compress("lznt1", host[blk << 20:(blk + 1) << 20].tobytes())
    ctx.close()


@pytest.mark.parametrize("fmt,each,count", [("xpress_huff", 65536, 16384), ("xpress", 1 << 20, 256)])
def test_batch_roundtrip_large(ms, oracle, fmt, each, count):
    import numpy as np
    import torch
    U = each * count
    host = np.empty(U, dtype=np.uint8)
    ms.generate_batch_into(host, each, count, 42, "text")
    ctx = ms.GpuContext()
    dev = ctx.device
    d_in = torch.from_numpy(host).to(dev)
    slot = (MAXSZ[fmt](each) + 15) // 16 * 16
    d_comp = torch.empty(slot * count, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(U, dtype=torch.uint8, device=dev)
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
    ctx.run_batch(fmt, "compress", d_in, d_comp, enc)
    ctx.run_batch(fmt, "decompress", d_comp, d_out, dec)
    torch.cuda.synchronize()
    assert int(st1.abs().sum()) == 0 and int(st2.abs().sum()) == 0
    assert bool((o_len == each).all())
    assert torch.equal(d_out, d_in)
    assert _checksum_of_checksums(d_out, 4096) == _checksum_of_checksums(d_in, 4096)
    clen = c_len.cpu().numpy()
    rnd = random.Random(2)
    for i in rnd.sample(range(count), 5):           # the reference's decoder accepts our streams
        c = d_comp[i * slot:i * slot + int(clen[i])].cpu().numpy().tobytes()
        assert oracle.decompress(fmt, c, each) == (0, host[i * each:(i + 1) * each].tobytes())
    ctx.close()
