"""LZNT1 parity on the B200, through the C ABI (lznt1_compress / lznt1_decompress / ms_*) and the device batch API.
Decode: bit-exact against the oracle on the same inputs.  Encode: byte-identical to the reference's encoder."""
import glob
import hashlib
import json
import os
import random

import pytest
from conftest import MAXSZ, golden_path

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 3, 7, 8, 31, 32, 33, 4095, 4096, 4097, 65535, 65536, 65537, 1 << 20]


def test_kat(ms):
    for kat in json.load(open(golden_path("kat.json"))):
        if kat["format"] != "lznt1":
            continue
        plain, comp = bytes.fromhex(kat["plain_hex"]), bytes.fromhex(kat["compressed_hex"])
        assert ms.raw_call("lznt1_decompress", comp, len(plain) + 10) == (0, plain)
        assert ms.raw_call("lznt1_compress", plain, len(plain) + 64) == (0, bytes.fromhex(kat["reference_encoder_hex"]))


@pytest.mark.parametrize("path", sorted(glob.glob(golden_path("ref_lznt1_*.bin"))), ids=os.path.basename)
def test_golden_fixture(ms, path):
    _, kind, n = os.path.basename(path)[4:-4].rsplit("_", 2)
    plain = ms.generate(kind, int(n), 42)
    comp = open(path, "rb").read()
    assert ms.raw_call("lznt1_decompress", comp, len(plain)) == (0, plain)
    assert ms.raw_call("lznt1_compress", plain, MAXSZ["lznt1"](len(plain))) == (0, comp)


@pytest.mark.parametrize("kind", ["text", "random", "runs"])
def test_decode_and_encode_parity(ms, oracle, kind):
    for n in SIZES:
        plain = ms.generate(kind, n, 42 + n % 11)
        comp = oracle.compress("lznt1", plain)
        assert ms.raw_call("lznt1_decompress", comp, n) == (0, plain), (kind, n)          # exact capacity
        assert ms.raw_call("lznt1_decompress", comp, n + 100) == (0, plain), (kind, n)
        assert ms.raw_call("lznt1_compress", plain, MAXSZ["lznt1"](n)) == (0, comp), (kind, n)
        assert ms.raw_call("lznt1_compress", plain, max(len(comp), 1)) == (0, comp), (kind, n)   # exact capacity
        if len(comp) > 1:
            assert ms.raw_call("lznt1_compress", plain, len(comp) - 1)[0] == ms.MSCOMP_BUF_ERROR


def test_generator_digests(ms):
    dig = json.load(open(golden_path("digests.json")))
    for key, e in dig.items():
        kind, n = key.rsplit("_", 1)
        plain = ms.generate(kind, int(n), 42)
        st, c = ms.raw_call("lznt1_compress", plain, MAXSZ["lznt1"](int(n)))
        assert st == 0 and len(c) == e["lznt1"]["size"] and hashlib.md5(c).hexdigest() == e["lznt1"]["md5"]


def test_degenerate_inputs(ms, oracle):
    for plain in (bytes(100000), b"ab" * 50000, bytes(range(256)) * 300, b"a" * 4096 + b"b" * 4096 + b"a" * 5000):
        comp = oracle.compress("lznt1", plain)
        assert ms.raw_call("lznt1_compress", plain, MAXSZ["lznt1"](len(plain))) == (0, comp)
        assert ms.raw_call("lznt1_decompress", comp, len(plain)) == (0, plain)


def test_heavy_chunks_are_byte_identical(ms, oracle):
    """Tiny alphabets / interleaved runs: huge hash buckets -> the encoder's lazy (search-on-visit) parse."""
    rnd = random.Random(11)
    cases = [bytes(rnd.choice(b"ACGT") for _ in range(64 * 4096 + 100)),
             b"".join(bytes([rnd.choice(b"\x00\x00\x00\xff\x01")]) * rnd.randint(1, 40) for _ in range(20000)),
             b"".join((b"\x00" * rnd.randint(3, 300)) + bytes([rnd.randint(1, 255)]) for _ in range(3000)),
             ms.generate("runs", 1 << 20, 5) + ms.generate("text", 1 << 18, 6)]
    for plain in cases:
        comp = oracle.compress("lznt1", plain)
        assert ms.raw_call("lznt1_compress", plain, MAXSZ["lznt1"](len(plain))) == (0, comp)
        assert ms.raw_call("lznt1_decompress", comp, len(plain)) == (0, plain)


def test_status_codes_match_oracle(ms, oracle):
    """SURVEY.md 8(b) table and bit flips: same status (and bytes on success) as the oracle."""
    rnd = random.Random(11)
    for kind in ("text", "runs"):
        n = 20000
        plain = ms.generate(kind, n, 9)
        comp = oracle.compress("lznt1", plain)
        cases = [(comp, cap) for cap in (n, n + 100, n - 1, n // 2, 0, 4 * n)]
        cases += [(comp[:-t], n) for t in (1, 2, 3, len(comp) // 2)]
        cases += [(comp + b"\0\0", n), (comp + b"\0\0", n + 10), (comp + b"\0", n), (comp + b"\x55", n),
                  (bytes([comp[0] ^ 0xFF, comp[1] ^ 0xFF]) + comp[2:], n)]
        for _ in range(60):
            b = bytearray(comp)
            for _ in range(rnd.randint(1, 3)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            cases.append((bytes(b), n + rnd.choice((0, 64))))
        for src, cap in cases:
            assert ms.raw_call("lznt1_decompress", src, cap) == oracle.decompress("lznt1", src, cap)


def test_short_chunks_in_mid_stream(ms, oracle):
    """Streams cut by a flushing compressor contain short chunks (src/lznt1_compress.cpp:162-166); decoder must compact."""
    a, b, c = ms.generate("text", 100, 1), ms.generate("text", 50, 2), ms.generate("text", 9000, 3)
    stream = b"".join(oracle.compress("lznt1", x) for x in (a, b, c))
    assert oracle.decompress("lznt1", stream, 20000) == (0, a + b + c)
    assert ms.raw_call("lznt1_decompress", stream, 20000) == (0, a + b + c)
    assert ms.raw_call("lznt1_decompress", stream, len(a + b + c)) == (0, a + b + c)


def test_ms_dispatch(ms, oracle):
    plain = ms.generate("text", 65536, 42)
    comp = ms.OpenSrc.LZNT1.Compress(plain)
    assert len(comp) == 40356                     # BASELINE.json config 1
    assert comp == oracle.compress("lznt1", plain)
    assert ms.OpenSrc.LZNT1.Decompress(comp, out_size=65536) == plain


def test_device_api_large_roundtrip(ms, oracle):
    """64 MiB through the device-resident API: round trip equals the input; sampled chunks equal the oracle's bytes."""
    import numpy as np
    import torch
    n = 64 << 20
    host = np.empty(n, dtype=np.uint8)
    ms.generate_batch_into(host, 1 << 20, n >> 20, 42, "text")
    ctx = ms.GpuContext()
    d_in = torch.from_numpy(host).cuda()
    d_comp, d_off, d_res = ctx.lznt1_compress(d_in)
    res = ms.parse_result(d_res)
    assert res["status"] == 0
    nchunks = n // 4096
    offs = d_off.cpu().numpy()
    assert int(offs[nchunks]) == res["bytes"]
    d_out, d_res2 = ctx.lznt1_decompress(d_comp, d_off, nchunks, in_len=res["bytes"])
    r2 = ms.parse_result(d_res2)
    assert r2 == {"bytes": n, "status": 0, "first_bad": 0, "short_units": 0}
    assert torch.equal(d_out[:n], d_in)
    comp = d_comp[:res["bytes"]].cpu().numpy().tobytes()
    for blk in (0, 17, 63):       # whole 1 MiB blocks against the oracle encoder
        lo, hi = int(offs[blk * 256]), int(offs[(blk + 1) * 256])
        assert comp[lo:hi] == oracle.compress("lznt1", host[blk << 20:(blk + 1) << 20].tobytes())
    assert ctx.launches() >= 5
    ctx.close()
