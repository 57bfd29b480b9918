"""Xpress and Xpress-Huffman parity on the B200 through the C ABI and the device batch API.
Decode: bit-exact against the oracle, same status codes.  Encode, both formats: the very bytes of the reference's encoders
(oracle == compiled reference, tests/test_oracle.py; reference-made digests in tests/golden/digests.json), and of course
round trips through the oracle's decoder."""
import hashlib
import glob
import json
import os
import random

import pytest
from conftest import MAXSZ, golden_path

pytestmark = pytest.mark.gpu
FMTS = ["xpress_huff", "xpress"]
SIZES = [0, 1, 2, 3, 7, 8, 31, 32, 33, 4095, 4096, 4097, 65535, 65536, 65537, 300000]


@pytest.mark.parametrize("fmt", FMTS)
def test_kat(ms, fmt):
    for kat in json.load(open(golden_path("kat.json"))):
        if kat["format"] != fmt:
            continue
        plain = bytes.fromhex(kat["plain_hex"])
        for comp in (bytes.fromhex(kat["compressed_hex"]), bytes.fromhex(kat["reference_encoder_hex"])):
            assert ms.raw_call(fmt + "_decompress", comp, len(plain)) == (0, plain)


@pytest.mark.parametrize("fmt", FMTS)
def test_golden_fixtures(ms, fmt):
    for path in sorted(glob.glob(golden_path("ref_%s_*.bin" % fmt))):
        if fmt == "xpress" and "xpress_huff" in path:
            continue
        _, kind, n = os.path.basename(path)[4:-4].rsplit("_", 2)
        plain = ms.generate(kind, int(n), 42)
        assert ms.raw_call(fmt + "_decompress", open(path, "rb").read(), len(plain)) == (0, plain), path


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("kind", ["text", "random", "runs"])
def test_decode_parity_and_encode_roundtrip(ms, oracle, fmt, kind):
    for n in SIZES:
        plain = ms.generate(kind, n, 42 + n % 11)
        ref_comp = oracle.compress(fmt, plain)
        if not (fmt == "xpress" and n == 0):       # upstream rejects its own empty Xpress stream
            assert ms.raw_call(fmt + "_decompress", ref_comp, n) == oracle.decompress(fmt, ref_comp, n) == (0, plain), (kind, n)
            assert ms.raw_call(fmt + "_decompress", ref_comp, n + 100) == (0, plain), (kind, n)
        st, comp = ms.raw_call(fmt + "_compress", plain, MAXSZ[fmt](n))
        assert st == 0 and len(comp) <= MAXSZ[fmt](n), (kind, n, st)
        if fmt == "xpress" and n == 0:
            assert comp == ref_comp == b"\xff\xff\xff\xff"
            continue
        assert comp == ref_comp, (fmt, kind, n)                               # byte-identical to the reference's encoder
        assert oracle.decompress(fmt, comp, n) == (0, plain), (kind, n)       # the reference's decoder accepts our stream
        assert ms.raw_call(fmt + "_decompress", comp, n) == (0, plain), (kind, n)
        assert ms.raw_call(fmt + "_compress", plain, len(comp)) == (0, comp)      # exact capacity
        if len(comp) > 1:
            assert ms.raw_call(fmt + "_compress", plain, len(comp) - 1)[0] == ms.MSCOMP_BUF_ERROR


@pytest.mark.parametrize("fmt", FMTS)
def test_status_codes_match_oracle(ms, oracle, fmt):
    rnd = random.Random(13)
    for kind in ("text", "runs"):
        n = 20000
        plain = ms.generate(kind, n, 9)
        comp = oracle.compress(fmt, plain)
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
            assert ms.raw_call(fmt + "_decompress", src, cap) == oracle.decompress(fmt, src, cap)


@pytest.mark.parametrize("fmt", FMTS)
def test_degenerate_inputs(ms, oracle, fmt):
    for plain in (bytes(300000), b"ab" * 50000, bytes(range(256)) * 300, b"abc" * 100, bytes(range(97, 123))):
        st, comp = ms.raw_call(fmt + "_compress", plain, MAXSZ[fmt](len(plain)))
        assert st == 0
        assert oracle.decompress(fmt, comp, len(plain)) == (0, plain)
        ref_comp = oracle.compress(fmt, plain)
        assert comp == ref_comp, (fmt, len(plain))
        assert ms.raw_call(fmt + "_decompress", ref_comp, len(plain)) == (0, plain)


@pytest.mark.parametrize("fmt", FMTS)
def test_generator_digests(ms, fmt):
    """md5 of our streams == md5 of the compiled reference's (made in the build container by make_golden.py)."""
    dig = json.load(open(golden_path("digests.json")))
    for key, d in dig.items():
        kind, n = key.rsplit("_", 1)
        plain = ms.generate(kind, int(n), 42)
        assert hashlib.md5(plain).hexdigest() == d["input_md5"]
        st, comp = ms.raw_call(fmt + "_compress", plain, MAXSZ[fmt](len(plain)))
        assert st == 0 and len(comp) == d[fmt]["size"] and hashlib.md5(comp).hexdigest() == d[fmt]["md5"], (fmt, key)


def test_xhuff_cross_chunk_matches_and_fallback(ms, oracle):
    """One-shot streams of several chunks: matches reach into the previous chunk (window 65535) and are clipped at the chunk end;
    chunks that would expand take the literal-only path -- same bytes as the reference either way.  Batch API: independent buffers."""
    rnd = random.Random(7)
    cases = [ms.generate("text", 65536, 1) * 3 + b"xyz", ms.generate("text", 20000, 7) * 9, bytes(200000), b"abc" * 70000,
             bytes(rnd.getrandbits(8) for _ in range(150000)), bytes(rnd.getrandbits(8) for _ in range(70000)) + ms.generate("text", 100000, 5),
             bytes(min(255, int(rnd.expovariate(1 / 40.0))) for _ in range(140000)), ms.generate("runs", 300000, 4), ms.generate("text", 1 << 20, 11)]
    for plain in cases:
        st, comp = ms.raw_call("xpress_huff_compress", plain, MAXSZ["xpress_huff"](len(plain)))
        assert st == 0 and comp == oracle.compress("xpress_huff", plain), len(plain)


def test_xpress_long_matches_and_forced_literals(ms, oracle):
    """Matches longer than the finder's cap are measured by the parse; after a match that outruns the reference's 8 KiB dictionary
    fills the reference emits literals where it would otherwise match (src/xpress_compress.cpp:269) -- same bytes here."""
    rnd = random.Random(5)
    cases = [bytes(9000) + ms.generate("text", 30000, 1), bytes(30000) + b"x" + bytes(30000) + ms.generate("text", 1000, 2),
             bytes(300000) + ms.generate("text", 3000, 2) + bytes(70000) + b"q" + bytes(70000), ms.generate("text", 20000, 7) * 3,
             bytes(rnd.getrandbits(8) & 3 for _ in range(70000)), ms.generate("text", 5000, 3) * 300, bytes(3 << 20)]
    for plain in cases:
        st, comp = ms.raw_call("xpress_compress", plain, MAXSZ["xpress"](len(plain)))
        assert st == 0 and comp == oracle.compress("xpress", plain), len(plain)


@pytest.mark.parametrize("fmt", FMTS)
def test_ms_dispatch(ms, oracle, fmt):
    codec = {"xpress": ms.OpenSrc.Xpress, "xpress_huff": ms.OpenSrc.XpressHuffman}[fmt]
    plain = ms.generate("text", 100000, 3)
    comp = codec.Compress(plain)
    assert codec.Decompress(comp, out_size=len(plain)) == plain
    assert oracle.decompress(fmt, comp, len(plain)) == (0, plain)


@pytest.mark.parametrize("encoder", ["reference", "fast"])
@pytest.mark.parametrize("fmt", FMTS)
def test_device_batch_roundtrip(ms, oracle, fmt, encoder):
    """1024 independent buffers of mixed sizes through the device batch API: every buffer round-trips, sampled buffers
    also decode with the oracle (and, with the default encoder, ARE the reference's bytes), reference-compressed buffers decode
    with our kernels in the same batch call.  encoder "fast": the bulk-synchronous matcher (mscomp_gpu_set_encoder)."""
    import numpy as np
    import torch
    rnd = random.Random(21)
    count = 1024
    sizes = [rnd.choice((65536, 65536, 65536, 1000, 30000, 1, 200000 if fmt == "xpress" else 65536)) for _ in range(count)]
    kinds = [("text", "runs", "random")[i % 3] for i in range(count)]
    plains = [ms.generate(k, n, 42 + i) for i, (k, n) in enumerate(zip(kinds, sizes))]
    in_off = np.cumsum([0] + [(n + 15) // 16 * 16 for n in sizes[:-1]]).tolist()
    blob = np.zeros(in_off[-1] + sizes[-1] + 16, dtype=np.uint8)
    for p, o in zip(plains, in_off):
        blob[o:o + len(p)] = np.frombuffer(p, dtype=np.uint8)
    caps = [MAXSZ[fmt](n) for n in sizes]
    out_off = np.cumsum([0] + [(c + 15) // 16 * 16 for c in caps[:-1]]).tolist()
    ctx = ms.GpuContext()
    ctx.set_encoder(encoder)
    assert ctx.encoder() == encoder
    d_in = torch.from_numpy(blob).cuda()
    d_comp = torch.zeros(out_off[-1] + caps[-1] + 16, dtype=torch.uint8, device="cuda")
    batch, olen, stat, keep = ctx.make_batch(in_off, sizes, out_off, caps)
    ctx.run_batch(fmt, "compress", d_in, d_comp, batch)
    clen = olen.cpu().numpy()
    assert not stat.cpu().numpy().any()
    comp_host = d_comp.cpu().numpy()
    for i in range(0, count, 37):
        c = comp_host[out_off[i]:out_off[i] + clen[i]].tobytes()
        if fmt == "xpress" and sizes[i] == 0:
            continue
        assert oracle.decompress(fmt, c, sizes[i]) == (0, plains[i]), i
        if encoder == "reference":
            assert c == oracle.compress(fmt, plains[i]), i             # the batch API writes the reference's bytes too
        elif kinds[i] == "text" and sizes[i] >= 30000:
            assert len(c) <= 1.12 * len(oracle.compress(fmt, plains[i])), i
    # decode everything back (every 5th buffer replaced by the reference encoder's stream)
    comps = [comp_host[out_off[i]:out_off[i] + clen[i]].tobytes() for i in range(count)]
    for i in range(0, count, 5):
        comps[i] = oracle.compress(fmt, plains[i])
    c_off = np.cumsum([0] + [(len(c) + 15) // 16 * 16 for c in comps[:-1]]).tolist()
    cblob = np.zeros(c_off[-1] + len(comps[-1]) + 16, dtype=np.uint8)
    for c, o in zip(comps, c_off):
        cblob[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
    d_c = torch.from_numpy(cblob).cuda()
    d_out = torch.zeros(in_off[-1] + sizes[-1] + 16, dtype=torch.uint8, device="cuda")
    b2, olen2, stat2, keep2 = ctx.make_batch(c_off, [len(c) for c in comps], in_off, sizes)
    ctx.run_batch(fmt, "decompress", d_c, d_out, b2)
    assert not stat2.cpu().numpy().any()
    assert olen2.cpu().numpy().tolist() == sizes
    out_host = d_out.cpu().numpy()
    for i in range(count):
        assert out_host[in_off[i]:in_off[i] + sizes[i]].tobytes() == plains[i], i
    ctx.close()


def test_xpress_restart_index_on_gpu(ms, oracle):
    """Side-band restart index of plain Xpress through the device API: streams equal the unindexed ones and decode with the
    oracle; the per-segment decoder reproduces the inputs."""
    import numpy as np
    import torch
    ctx = ms.GpuContext()
    rnd = random.Random(4)
    sizes = [1, 2047, 2048, 2049, 70000, 300000, 1 << 20, (1 << 20) + 5, 0]
    plains = [ms.generate(("text", "runs", "random")[i % 3], n, 11 + i) for i, n in enumerate(sizes)]
    plains.append(bytes(rnd.getrandbits(8) & 3 for _ in range(100000)))
    count = len(plains)
    caps = [MAXSZ["xpress"](len(p)) for p in plains]
    in_off = np.cumsum([0] + [(len(p) + 63) // 64 * 64 for p in plains])[:-1]
    c_off = np.cumsum([0] + [(c + 63) // 64 * 64 for c in caps])[:-1]
    blob = np.zeros(int(in_off[-1]) + len(plains[-1]) + 64, dtype=np.uint8)
    for p, o in zip(plains, in_off):
        blob[o:o + len(p)] = np.frombuffer(p, dtype=np.uint8)
    d_in = torch.from_numpy(blob).cuda()
    d_comp = torch.zeros(int(c_off[-1]) + caps[-1] + 64, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros_like(d_in)
    enc, c_len, st1, keep1 = ctx.make_batch(in_off.tolist(), [len(p) for p in plains], c_off.tolist(), caps)
    for seg_bytes, stride in ((2048, 600), (16384, 65), (65536, 4)):
        d_index = ctx.xpress_index_tensor(count, stride)
        ctx.xpress_compress_indexed(d_in, d_comp, enc, seg_bytes, d_index, stride)
        torch.cuda.synchronize()
        assert int(st1.abs().sum().item()) == 0
        comps = [bytes(d_comp[o:o + n].cpu().numpy()) for o, n in zip(c_off.tolist(), c_len.tolist())]
        d_plain = torch.zeros_like(d_comp)
        enc2, c_len2, st1b, keep1b = ctx.make_batch(in_off.tolist(), [len(p) for p in plains], c_off.tolist(), caps)
        ctx.run_batch("xpress", "compress", d_in, d_plain, enc2)
        torch.cuda.synchronize()
        assert c_len2.tolist() == c_len.tolist()
        for p, c, o in zip(plains, comps, c_off.tolist()):
            assert bytes(d_plain[o:o + len(c)].cpu().numpy()) == c              # the index does not change the stream
            if p:
                assert oracle.decompress("xpress", c, len(p)) == (0, p)
        dec, o_len, st2, keep2 = ctx.make_batch(c_off.tolist(), c_len.tolist(), in_off.tolist(), [len(p) for p in plains])
        d_out.zero_()
        ctx.xpress_decompress_indexed(d_comp, d_out, dec, d_index, stride)
        torch.cuda.synchronize()
        assert st2.tolist() == [0 if p else -3 for p in plains]                 # (FF FF FF FF is rejected, as upstream)
        for p, o, n in zip(plains, in_off.tolist(), o_len.tolist()):
            if p:
                assert n == len(p) and bytes(d_out[o:o + n].cpu().numpy()) == p


def test_xhuff_restart_index_on_gpu(ms, oracle):
    """Xpress-Huffman through the restart index: the serial walk (statuses, lengths, index) followed by the one-lane-per-segment
    decoder must give the oracle's bytes and statuses -- for reference-made streams (multi-chunk, cross-chunk matches), damaged
    streams and a damaged index (which sends the stream to the index-free decoder instead of into wrong output)."""
    import numpy as np
    import torch
    rnd = random.Random(21)
    ctx = ms.GpuContext()
    plains = [ms.generate(k, n, 5 + n % 7) for k in ("text", "runs", "random") for n in (1, 100, 5000, 65536, 70000, 200000)]
    plains += [b"abc" * 30000, bytes(150000), b""]
    comps = [oracle.compress("xpress_huff", p) for p in plains]
    caps = [len(p) for p in plains]
    # damaged streams: the walk is the authority on their status
    base = oracle.compress("xpress_huff", ms.generate("text", 20000, 9))
    for _ in range(40):
        b = bytearray(base)
        for _ in range(rnd.randint(1, 3)):
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        comps.append(bytes(b)); caps.append(20000 + rnd.choice((0, 64)))
    comps += [base[:-1], base[:len(base) // 2], base + b"\0\0"]; caps += [20000, 20000, 20000]
    c_off = np.cumsum([0] + [(len(c) + 31) // 16 * 16 for c in comps])
    o_off = np.cumsum([0] + [(c + 31) // 16 * 16 for c in caps])
    blob = np.zeros(int(c_off[-1]) + 64, dtype=np.uint8)
    for c, o in zip(comps, c_off):
        blob[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
    d_c = torch.from_numpy(blob).to(ctx.device)
    d_out = torch.zeros(int(o_off[-1]) + 64, dtype=torch.uint8, device=ctx.device)
    dec, o_len, st, keep = ctx.make_batch(c_off[:-1].tolist(), [len(c) for c in comps], o_off[:-1].tolist(), caps)
    d_index, d_ioff, d_nseg = ctx.xhuff_index_alloc(caps)
    before = ctx.xhuff_second_chances()
    ctx.xhuff_index(d_c, dec, d_index, d_ioff, d_nseg)
    walk_st, walk_len = st.cpu().tolist(), o_len.cpu().tolist()
    ctx.xhuff_decompress_indexed(d_c, d_out, dec, d_index, d_ioff, d_nseg)
    ctx.synchronize()
    out = d_out.cpu().numpy()
    got_st, got_len = st.cpu().tolist(), o_len.cpu().tolist()
    for i, (c, cap) in enumerate(zip(comps, caps)):
        rc, exp = oracle.decompress("xpress_huff", c, cap)
        assert walk_st[i] == rc and got_st[i] == rc, (i, walk_st[i], got_st[i], rc)
        if rc == 0:
            assert walk_len[i] == got_len[i] == len(exp)
            assert out[o_off[i]:o_off[i] + got_len[i]].tobytes() == exp, i
    # valid streams never need the index-free decoder (the damaged ones that the walk rejects reach it through the trusted entry point)
    nvalid = len(plains)
    assert ctx.xhuff_second_chances() - before <= len(comps) - nvalid
    # a wrong index: shift the byte cursor of one entry of every stream that has at least 3
    idx = d_index.cpu().numpy().reshape(-1, 4).copy()
    ioff, nseg = d_ioff.cpu().numpy(), d_nseg.cpu().numpy()
    hit = 0
    for i in range(nvalid):
        if nseg[i] >= 3:
            idx[ioff[i] + 2, 1] += 2
            hit += 1
    d_bad = torch.from_numpy(idx.reshape(-1)).to(ctx.device)
    d_out.zero_()
    before = ctx.xhuff_second_chances()
    ctx.xhuff_decompress_indexed(d_c, d_out, dec, d_bad, d_ioff, d_nseg)
    ctx.synchronize()
    out = d_out.cpu().numpy()
    got_st, got_len = st.cpu().tolist(), o_len.cpu().tolist()
    for i in range(nvalid):
        assert got_st[i] == 0 and out[o_off[i]:o_off[i] + got_len[i]].tobytes() == plains[i], i
    assert ctx.xhuff_second_chances() - before >= hit


@pytest.mark.gpu
def test_xhuff_encoder_index_on_gpu(ms, oracle):
    """mscomp_gpu_xpress_huff_compress_batch_indexed: same streams as the plain call (the oracle decodes them), and an index that
    equals the serial walk's entry for entry; the indexed decoder gives the plain bytes back without a single second chance."""
    import numpy as np
    import torch
    ctx = ms.GpuContext()
    plains = [ms.generate(k, n, 3 + n % 5) for k in ("text", "runs", "random") for n in (1, 63, 4096, 65536, 65537, 200000)] + [b"ab" * 70000, bytes(100000)]
    caps = [len(p) + 34 + 258 + 258 * (len(p) // 65536) + 16 for p in plains]
    i_off = np.cumsum([0] + [(len(p) + 31) // 16 * 16 for p in plains])
    c_off = np.cumsum([0] + [(c + 31) // 16 * 16 for c in caps])
    blob = np.zeros(int(i_off[-1]) + 64, dtype=np.uint8)
    for p, o in zip(plains, i_off):
        blob[o:o + len(p)] = np.frombuffer(p, dtype=np.uint8)
    d_in = torch.from_numpy(blob).to(ctx.device)
    d_c = torch.zeros(int(c_off[-1]) + 64, dtype=torch.uint8, device=ctx.device)
    d_c2 = torch.zeros_like(d_c)
    enc, c_len, st, keep = ctx.make_batch(i_off[:-1].tolist(), [len(p) for p in plains], c_off[:-1].tolist(), caps)
    d_index, d_ioff, d_nseg = ctx.xhuff_index_alloc([len(p) for p in plains])
    ctx.xhuff_compress_indexed(d_in, d_c, enc, d_index, d_ioff, d_nseg)
    ctx.synchronize()
    assert not st.cpu().numpy().any()
    lens = c_len.cpu().tolist()
    mine, n_mine = d_index.cpu().numpy().reshape(-1, 4).copy(), d_nseg.cpu().numpy().copy()
    ctx.run_batch("xpress_huff", "compress", d_in, d_c2, enc)
    ctx.synchronize()
    assert c_len.cpu().tolist() == lens and torch.equal(d_c, d_c2)
    comp = d_c.cpu().numpy()
    for i, p in enumerate(plains):
        rc, back = oracle.decompress("xpress_huff", comp[c_off[i]:c_off[i] + lens[i]].tobytes(), len(p))
        assert rc == 0 and back == p, i
    # the walk over the same streams
    dec, o_len, st2, keep2 = ctx.make_batch(c_off[:-1].tolist(), lens, i_off[:-1].tolist(), [len(p) for p in plains])
    d_index2, d_ioff2, d_nseg2 = ctx.xhuff_index_alloc([len(p) for p in plains])
    ctx.xhuff_index(d_c, dec, d_index2, d_ioff2, d_nseg2)
    ctx.synchronize()
    walk, n_walk = d_index2.cpu().numpy().reshape(-1, 4), d_nseg2.cpu().numpy()
    ioff = d_ioff.cpu().numpy()
    assert (n_mine == n_walk).all()
    for i in range(len(plains)):
        assert (mine[ioff[i]:ioff[i] + n_mine[i]] == walk[ioff[i]:ioff[i] + n_walk[i]]).all(), i
    # decode from the encoder's index
    d_out = torch.zeros(int(i_off[-1]) + 64, dtype=torch.uint8, device=ctx.device)
    before = ctx.xhuff_second_chances()
    ctx.xhuff_decompress_indexed(d_c, d_out, dec, d_index, d_ioff, d_nseg)
    ctx.synchronize()
    assert not st2.cpu().numpy().any() and o_len.cpu().tolist() == [len(p) for p in plains]
    out = d_out.cpu().numpy()
    for i, p in enumerate(plains):
        assert out[i_off[i]:i_off[i] + len(p)].tobytes() == p, i
    assert ctx.xhuff_second_chances() == before


@pytest.mark.gpu
def test_xhuff_one_shot_multichunk_streams(ms, oracle):
    """xpress_huff_decompress on ONE multi-chunk stream takes the chunk-parallel path (kernels/xhuff_chunks.cuh): reference-made
    streams (matches reach into the previous chunk), streams made here (independent chunks), and everything that path must leave to
    the serial decoder -- too little room, damaged tables and payloads, truncation, trailing bytes -- with the oracle's status."""
    rnd = random.Random(5)
    for kind, n in (("text", 1 << 23), ("text", 300001), ("runs", 1 << 20), ("random", 200000), ("text", 65536 * 2)):
        plain = ms.generate(kind, n, 3 + n % 7)
        ref = oracle.compress("xpress_huff", plain)
        assert ms.raw_call("xpress_huff_decompress", ref, n) == (0, plain), (kind, n)
        assert ms.raw_call("xpress_huff_decompress", ref, n + 1000) == (0, plain), (kind, n)
        st, mine = ms.raw_call("xpress_huff_compress", plain, MAXSZ["xpress_huff"](n))
        assert st == 0 and ms.raw_call("xpress_huff_decompress", mine, n) == (0, plain), (kind, n)
        assert oracle.decompress("xpress_huff", mine, n) == (0, plain)
        if n > 400000:
            continue
        cases = [(ref, n - 1), (ref, n // 2), (ref, 0), (ref[:-1], n), (ref[:-2], n), (ref[:len(ref) // 2], n), (ref + bytes(2), n + 10), (ref + b"U", n + 10)]
        for _ in range(25):
            b = bytearray(ref)
            for _ in range(rnd.randint(1, 3)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            cases.append((bytes(b), n + rnd.choice((0, 64))))
        for comp, cap in cases:
            exp = oracle.decompress("xpress_huff", comp, cap)
            got = ms.raw_call("xpress_huff_decompress", comp, cap)
            assert got[0] == exp[0] and (exp[0] != 0 or got[1] == exp[1]), (kind, n, cap, got[0], exp[0])


@pytest.mark.gpu
def test_xpress_one_shot_single_stream(ms, oracle):
    """xpress_decompress on ONE stream of some size takes the speculative path (kernels/xpress_spec.cuh: the flag group that would
    start at every offset, the true chain linked up, a restart index from it, indexed parse + order-free token execution):
    reference-made streams of every kind, streams made here, and everything that path must leave to the serial decoder -- too little
    room, damage, truncation, trailing bytes -- with the oracle's status."""
    rnd = random.Random(7)
    plains = [ms.generate(kind, n, 3 + n % 7) for kind, n in (("text", 1 << 23), ("text", 300001), ("runs", 1 << 22), ("random", 200000), ("text", 90000))]
    plains += [bytes(rnd.randrange(4) for _ in range(300000)),                       # long matches everywhere: busy length nibbles
               b"".join(bytes([rnd.randrange(256)]) * rnd.choice((1, 2, 9, 10, 11, 24, 25, 26, 300, 70000)) for _ in range(20000))]
    for plain in plains:
        n = len(plain)
        ref = oracle.compress("xpress", plain)
        assert ms.raw_call("xpress_decompress", ref, n) == (0, plain), n
        assert ms.raw_call("xpress_decompress", ref, n + 1000) == (0, plain), n
        st, mine = ms.raw_call("xpress_compress", plain, MAXSZ["xpress"](n))
        assert st == 0 and mine == ref and ms.raw_call("xpress_decompress", mine, n) == (0, plain), n
        if n > 400000 or len(ref) < 32768:
            continue
        cases = [(ref, n - 1), (ref, n // 2), (ref, 0), (ref[:-1], n), (ref[:-2], n), (ref[:len(ref) // 2], n), (ref + bytes(2), n + 10), (ref + b"U", n + 10)]
        for _ in range(25):
            b = bytearray(ref)
            for _ in range(rnd.randint(1, 3)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            cases.append((bytes(b), n + rnd.choice((0, 64))))
        for comp, cap in cases:
            exp = oracle.decompress("xpress", comp, cap)
            got = ms.raw_call("xpress_decompress", comp, cap)
            assert got[0] == exp[0] and (exp[0] != 0 or got[1] == exp[1]), (n, cap, got[0], exp[0])
