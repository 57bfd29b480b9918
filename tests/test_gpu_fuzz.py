"""Seeded structural fuzz on the B200: a few hundred small buffers built from random bytes, repeated segments, short
periods and long runs go through every codec in one batch call per direction and are cross-checked against the oracle
in both directions (our encoder -> reference decoder, reference encoder -> our decoder)."""
import random

import pytest
from conftest import MAXSZ

pytestmark = pytest.mark.gpu


def make_buffers(seed, count):
    rnd = random.Random(seed)
    bufs = []
    for _ in range(count):
        n = rnd.choice((0, 1, 2, 3, 5, 17, 100, 1000, 4096, 4097, 9000, 20000, 70000))
        out = bytearray()
        while len(out) < n:
            kind = rnd.randrange(5)
            if kind == 0:
                out += bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 40)))
            elif kind == 1 and out:
                off = rnd.randint(1, min(len(out), 9000))
                ln = rnd.randint(3, 300)
                for _ in range(ln):
                    out.append(out[-off])
            elif kind == 2:
                out += bytes([rnd.randrange(256)]) * rnd.randint(1, 600)
            elif kind == 3:
                pat = bytes(rnd.randrange(256) for _ in range(rnd.randint(2, 7)))
                out += pat * rnd.randint(1, 80)
            else:
                out += bytes(rnd.choice(b"etaoin shrdlu") for _ in range(rnd.randint(1, 60)))
        bufs.append(bytes(out[:n]))
    return bufs


def run_batch(ms, ctx, fmt, op, bufs, caps):
    import numpy as np
    import torch
    in_off, pos = [], 0
    for b in bufs:
        in_off.append(pos)
        pos += (len(b) + 15) // 16 * 16 + 16
    blob = np.zeros(pos + 16, dtype=np.uint8)
    for b, o in zip(bufs, in_off):
        blob[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    out_off, opos = [], 0
    for c in caps:
        out_off.append(opos)
        opos += (c + 15) // 16 * 16 + 16
    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(opos + 16, dtype=torch.uint8, device="cuda")
    batch, olen, stat, keep = ctx.make_batch(in_off, [len(b) for b in bufs], out_off, caps)
    ctx.run_batch(fmt, op, d_in, d_out, batch)
    ol, st, host = olen.cpu().numpy(), stat.cpu().numpy(), d_out.cpu().numpy()
    return st, [host[out_off[i]:out_off[i] + ol[i]].tobytes() for i in range(len(bufs))]


@pytest.mark.parametrize("fmt", ["xpress_huff", "xpress"])
def test_batch_fuzz_both_directions(ms, oracle, fmt):
    bufs = make_buffers(1234, 400)
    ctx = ms.GpuContext()
    st, comps = run_batch(ms, ctx, fmt, "compress", bufs, [MAXSZ[fmt](len(b)) for b in bufs])
    assert not st.any()
    for b, c in zip(bufs, comps):
        assert len(c) <= MAXSZ[fmt](len(b))
        if fmt == "xpress" and not b:
            continue
        assert oracle.decompress(fmt, c, len(b)) == (0, b)
    ref = [oracle.compress(fmt, b) for b in bufs]
    keep = [i for i, b in enumerate(bufs) if not (fmt == "xpress" and not b)]
    st, outs = run_batch(ms, ctx, fmt, "decompress", [ref[i] for i in keep] + [comps[i] for i in keep],
                         [len(bufs[i]) for i in keep] * 2)
    assert not st.any()
    assert outs == [bufs[i] for i in keep] * 2
    ctx.close()


def test_lznt1_fuzz_byte_identical(ms, oracle):
    for b in make_buffers(99, 120):
        if not b:
            continue
        comp = oracle.compress("lznt1", b)
        assert ms.raw_call("lznt1_compress", b, MAXSZ["lznt1"](len(b))) == (0, comp)
        assert ms.raw_call("lznt1_decompress", comp, len(b)) == (0, b)
