"""Xpress / Xpress-Huffman CUDA kernels compiled for the CPU (tests/emu) against the oracle.
Decoders: bit-exact output and identical status.  Encoders: the stream must decode with the oracle (= the reference's
decoder restated) back to the input, stay inside the reference's max_compressed_size bound and near the reference's ratio."""
import random

import pytest
from conftest import MAXSZ
import ctypes

import numpy as np
from emu_util import emu_xhuff_stream_exact, P, PAD, emu_batch, padded

KERNELS = {"xpress_huff": ("emu_xhuff_decode", "emu_xhuff_encode"), "xpress": ("emu_xpress_decode", "emu_xpress_encode")}
SIZES = [1, 2, 3, 4, 5, 31, 32, 33, 100, 5000, 65535, 65536, 65537, 140000]


def corpus(ms):
    out = []
    for kind in ("text", "runs", "random"):
        out += [ms.generate(kind, n, 42 + n) for n in SIZES]
    out += [b"abc" * 100, bytes(100000), bytes(range(97, 123)), b"ab" * 40000, bytes(range(256)) * 300]
    return out


DECODERS = {"xpress_huff": ["emu_xhuff_decode", "emu_xhuff_decode2", "emu_xhuff_decode2"],           # fused, parse+apply
            "xpress": ["emu_xpress_decode", "emu_xpress_decode2", "emu_xpress_decode_warp"]}         # fused, lane parse + warp apply, warp parse + block apply


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("fmt", ["xpress_huff", "xpress"])
def test_decode_matches_oracle(emu, oracle, ms, fmt, variant):
    plains = corpus(ms)
    comps = [oracle.compress(fmt, p) for p in plains]
    st, ol, outs = emu_batch(emu, DECODERS[fmt][variant], comps + [b""], [len(p) for p in plains] + [0])
    assert not st.any()
    assert outs[:-1] == plains and outs[-1] == b""


@pytest.mark.parametrize("fmt", ["xpress_huff", "xpress"])
def test_encode_roundtrips_through_oracle(emu, oracle, ms, fmt):
    plains = corpus(ms) + [b""]
    caps = [MAXSZ[fmt](len(p)) for p in plains]
    st, ol, outs = emu_batch(emu, KERNELS[fmt][1], plains, caps)
    assert not st.any()
    for p, c in zip(plains, outs):
        assert len(c) <= MAXSZ[fmt](len(p))
        if fmt == "xpress" and not p:
            assert c == b"\xff\xff\xff\xff"          # upstream cannot decode its own empty stream
            continue
        assert oracle.decompress(fmt, c, len(p)) == (0, p)
    # ratio: within 15 % of the reference encoder on the 64 KiB text buffer
    p = ms.generate("text", 65536, 42 + 65536)
    ours = outs[plains.index(p)]
    assert len(ours) <= 1.15 * len(oracle.compress(fmt, p))


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("fmt", ["xpress_huff", "xpress"])
def test_decode_status_equals_oracle_on_bad_input(emu, oracle, ms, fmt, variant):
    rnd = random.Random(5)
    cases = []
    for kind in ("text", "runs"):
        n = 20000
        plain = ms.generate(kind, n, 9)
        comp = oracle.compress(fmt, plain)
        cases += [(comp, cap) for cap in (n, n + 100, n - 1, n // 2, 0)]
        cases += [(comp[:-t], n) for t in (1, 2, 3, len(comp) // 2)]
        cases += [(comp + b"\0\0", n), (comp + b"\0\0", n + 10), (comp + b"\x55", n)]
        for _ in range(40):
            b = bytearray(comp)
            for _ in range(rnd.randint(1, 3)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            cases.append((bytes(b), n + rnd.choice((0, 64))))
    st, ol, outs = emu_batch(emu, DECODERS[fmt][variant], [c for c, _ in cases], [cap for _, cap in cases])
    for i, (src, cap) in enumerate(cases):
        rc, exp = oracle.decompress(fmt, src, cap)
        assert st[i] == rc, (i, st[i], rc)
        if rc == 0:
            assert outs[i] == exp


@pytest.mark.parametrize("block_apply", [0, 1], ids=["warp-per-stream", "block-per-stream"])
@pytest.mark.parametrize("seg_bytes,stride", [(2048, 80), (4096, 8), (16384, 16)])
def test_xpress_restart_index(emu, oracle, ms, seg_bytes, stride, block_apply):
    """The side-band restart index of plain Xpress: the stream itself is unchanged (same bytes as the plain call, decodes with
    the oracle), and the one-lane-per-segment decoder reproduces the input from (stream, index)."""
    rnd = random.Random(seg_bytes)
    plains = [ms.generate(kind, n, 7 + n) for kind in ("text", "runs", "random") for n in (1, 2047, 2048, 2049, 10000, 70000, 140000)]
    plains += [b"", b"abc" * 30000, bytes(50000), bytes(rnd.getrandbits(8) & 3 for _ in range(30000))]
    caps = [MAXSZ["xpress"](len(p)) for p in plains]
    index = np.zeros(len(plains) * stride * 6, dtype=np.uint32)
    st, _, comps = emu_batch(emu, "emu_xpress_encode_indexed", plains, caps,
                             extra=(ctypes.c_uint32(seg_bytes), P(index), ctypes.c_uint32(stride)))
    assert not st.any()
    st0, _, comps0 = emu_batch(emu, "emu_xpress_encode", plains, caps)
    assert comps == comps0                                           # the index does not change the stream
    for p, c in zip(plains, comps):
        if p:                                                        # (the reference rejects its own encoding of the empty input)
            assert oracle.decompress("xpress", c, len(p)) == (0, p)
    ix = index.reshape(len(plains), stride, 6)
    for i, p in enumerate(plains):
        want = min(stride, max(1, -(-len(p) // seg_bytes)))
        assert ix[i, 0, 5] == want, (len(p), ix[i, 0])
        outs_pos = [int(x) for x in ix[i, :want, 1]]
        assert outs_pos == sorted(outs_pos) and outs_pos[0] == 0
    st, ol, outs = emu_batch(emu, "emu_xpress_decode_indexed", comps, [len(p) for p in plains], extra=(P(index), ctypes.c_uint32(stride), ctypes.c_uint32(block_apply)))
    st_plain, _, outs_plain = emu_batch(emu, "emu_xpress_decode2", comps, [len(p) for p in plains])
    assert list(st) == list(st_plain) and list(st) == [0 if p else -3 for p in plains]       # (FF FF FF FF is rejected, as upstream)
    assert outs == outs_plain and all(o == p for o, p in zip(outs, plains) if p)
    # exact / too small capacities and damaged streams: the status of the plain decoder (which equals the oracle's), no crash
    for i in rnd.sample(range(len(plains)), 6):
        if len(plains[i]) < 100:
            continue
        bad = bytearray(comps[i])
        bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        one_ix = ix[i:i + 1].copy().reshape(-1)
        for data, cap in ((bytes(bad), len(plains[i])), (comps[i], len(plains[i]) - 1), (comps[i], len(plains[i]) // 2)):
            st1, _, out1 = emu_batch(emu, "emu_xpress_decode_indexed", [data], [cap], extra=(P(one_ix), ctypes.c_uint32(stride), ctypes.c_uint32(block_apply)))
            ref = oracle.decompress("xpress", data, cap)
            if data is comps[i]:
                assert st1[0] == ref[0] and (ref[0] < 0 or out1[0] == ref[1])
            else:                                                    # the index no longer belongs to the stream: an error, or the right bytes
                assert st1[0] < 0 or (ref[0] == 0 and out1[0] == ref[1])


@pytest.mark.parametrize("with_index", [False, True])
def test_xpress_two_pass_encoder(emu, oracle, ms, with_index):
    """Two-pass plain Xpress encoder (match finding per 64 KiB segment, emission per buffer): every stream decodes with the
    oracle, stays inside the reference's bound and close to the single-pass encoder's size; the restart index it writes works."""
    rnd = random.Random(9)
    plains = [ms.generate(kind, n, 3 + n) for kind in ("text", "runs", "random") for n in (5, 65537, 140000)]
    plains += [ms.generate("text", 1, 1), ms.generate("text", 65536, 2), ms.generate("text", 300000, 3)]
    plains += [b"", b"abc" * 50000, bytes(140000), bytes(rnd.getrandbits(8) & 3 for _ in range(70000))]
    caps = [MAXSZ["xpress"](len(p)) for p in plains]
    seg_bytes, stride = 16384, 24
    index = np.zeros(len(plains) * stride * 6, dtype=np.uint32)
    extra = (ctypes.c_uint32(seg_bytes), P(index) if with_index else ctypes.c_void_p(0), ctypes.c_uint32(stride))
    st, _, comps = emu_batch(emu, "emu_xpress_encode2", plains, caps, extra=extra)
    assert not st.any()
    st1, _, comps1 = emu_batch(emu, "emu_xpress_encode", plains, caps)
    for p, c, c1 in zip(plains, comps, comps1):
        assert len(c) <= MAXSZ["xpress"](len(p))
        if p:
            assert oracle.decompress("xpress", c, len(p)) == (0, p)
            assert len(c) <= len(c1) * 1.02 + 16, (len(p), len(c), len(c1))       # segment boundaries cost next to nothing
    # exact-capacity and too-small outputs (with an index array of their own)
    index2 = np.zeros(len(plains) * stride * 6, dtype=np.uint32)
    extra2 = (ctypes.c_uint32(seg_bytes), P(index2) if with_index else ctypes.c_void_p(0), ctypes.c_uint32(stride))
    st2, _, _ = emu_batch(emu, "emu_xpress_encode2", plains, [len(c) for c in comps], extra=extra2)
    assert not st2.any()
    st3, _, _ = emu_batch(emu, "emu_xpress_encode2", [p for p in plains if len(p) > 100], [len(c) - 1 for p, c in zip(plains, comps) if len(p) > 100], extra=extra2)
    assert (st3 == -5).all()
    if with_index:
        for blk in (0, 1):
            st4, _, outs = emu_batch(emu, "emu_xpress_decode_indexed", comps, [len(p) for p in plains],
                                     extra=(P(index), ctypes.c_uint32(stride), ctypes.c_uint32(blk)))
            assert list(st4) == [0 if p else -3 for p in plains]
            assert all(o == p for o, p in zip(outs, plains) if p)


def test_xpress_reference_parse_is_byte_identical(emu, oracle, ms):
    """The default plain-Xpress encoder (lz_exact.cuh: hash-chain links, Find at every position, greedy parse + emission) writes the
    very bytes of the reference's xpress_compress: edge sizes, the end-of-input length quirk, hash collisions (4-symbol data), matches
    longer than LZX_CAP that the parse extends, and the literals the reference is forced into after a match that outruns its 8 KiB
    dictionary fills (src/xpress_compress.cpp:269)."""
    rnd = random.Random(5)
    plains = [ms.generate(kind, n, 3 + n) for kind in ("text", "runs", "random") for n in (1, 2, 3, 4, 5, 31, 32, 33, 100, 4096, 70000)]
    plains += [b"", b"abc" * 50000, bytes(140000), bytes(rnd.getrandbits(8) & 3 for _ in range(70000)),
               bytes(9000) + ms.generate("text", 30000, 1), bytes(30000) + b"x" + bytes(30000) + ms.generate("text", 1000, 2),
               bytes(300000) + ms.generate("text", 3000, 2) + bytes(70000) + b"q" + bytes(70000),
               ms.generate("text", 20000, 7) * 3, ms.generate("text", 140001, 9)]
    caps = [MAXSZ["xpress"](len(p)) for p in plains]
    seg_bytes, stride = 16384, 24
    index = np.zeros(len(plains) * stride * 6, dtype=np.uint32)
    st, _, comps = emu_batch(emu, "emu_xpress_encode_exact", plains, caps, grid=3, extra=(ctypes.c_uint32(seg_bytes), P(index), ctypes.c_uint32(stride)))
    assert not st.any()
    for p, c in zip(plains, comps):
        assert c == oracle.compress("xpress", p), len(p)
    # exact capacity passes, one byte less is a buffer error
    big = [(p, c) for p, c in zip(plains, comps) if len(p) > 100]
    st2, _, _ = emu_batch(emu, "emu_xpress_encode_exact", [p for p, _ in big], [len(c) for _, c in big],
                          extra=(ctypes.c_uint32(seg_bytes), ctypes.c_void_p(0), ctypes.c_uint32(0)))
    assert not st2.any()
    st3, _, _ = emu_batch(emu, "emu_xpress_encode_exact", [p for p, _ in big], [len(c) - 1 for _, c in big],
                          extra=(ctypes.c_uint32(seg_bytes), ctypes.c_void_p(0), ctypes.c_uint32(0)))
    assert (st3 == -5).all()
    # the restart index written along the way works
    for blk in (0, 1):
        st4, _, outs = emu_batch(emu, "emu_xpress_decode_indexed", comps, [len(p) for p in plains],
                                 extra=(P(index), ctypes.c_uint32(stride), ctypes.c_uint32(blk)))
        assert list(st4) == [0 if p else -3 for p in plains]
        assert all(o == p for o, p in zip(outs, plains) if p)


def test_xhuff_reference_encoder_is_byte_identical(emu, oracle, ms):
    """The default Xpress-Huffman encoder writes the very bytes of the reference's xpress_huff_compress: its LZ77 tokens (lz_exact.cuh,
    window 65535 across chunks, clipped at the chunk), the code lengths of its heap-driven CreateCodes replayed tie for tie, and the
    literal-only fallback on chunks that would expand.  Independent buffers (one or several chunks each) and the one-shot layout (the
    chunks of ONE stream as the units of the batch, matches reaching into the previous chunk)."""
    rnd = random.Random(7)
    plains = [ms.generate(kind, n, 3 + n) for kind in ("text", "runs", "random") for n in (1, 2, 3, 5, 100, 4096, 65535, 65536, 70000)]
    plains += [b"abc" * 30000, bytes(140000), bytes(rnd.getrandbits(8) & 3 for _ in range(70000)), ms.generate("text", 20000, 7) * 8,
               ms.generate("text", 200001, 9), ms.generate("text", 65536, 1) * 2 + b"xyz",
               bytes(min(255, int(rnd.expovariate(1 / 40.0))) for _ in range(100000)),
               bytes(rnd.getrandbits(8) for _ in range(40000)) + ms.generate("text", 60000, 5)]
    st, _, comps = emu_batch(emu, "emu_xhuff_encode_exact", plains, [MAXSZ["xpress_huff"](len(p)) for p in plains], grid=3, extra=(ctypes.c_uint32(0),))
    assert not st.any()
    for p, c in zip(plains, comps):
        assert c == oracle.compress("xpress_huff", p), len(p)
    for p in plains:
        if len(p) > 65536:
            st1, c1 = emu_xhuff_stream_exact(emu, p)
            assert not st1.any() and c1 == oracle.compress("xpress_huff", p), len(p)


# ---- Xpress-Huffman through the restart index (xhuff_index.cuh + xhuff_decode_indexed.cuh) -----------------------------------

def _second(emu):
    emu.emu_xhuff_second_chances.restype = ctypes.c_uint64
    return int(emu.emu_xhuff_second_chances())


@pytest.mark.parametrize("trusted", [0, 1])
def test_xhuff_indexed_decode_matches_oracle(emu, oracle, ms, trusted):
    """Serial walk -> index -> one lane per segment.  trusted=1: the decoder takes lengths / statuses from the index alone (the public
    indexed entry point); valid streams must never need the second-chance (index-free) decoder."""
    plains = corpus(ms)
    comps = [oracle.compress("xpress_huff", p) for p in plains]
    before = _second(emu)
    st, ol, outs = emu_batch(emu, "emu_xhuff_decode_indexed", comps + [b""], [len(p) for p in plains] + [0],
                             extra=(ctypes.c_uint32(0), ctypes.c_uint32(trusted)))
    assert not st.any()
    assert outs[:-1] == plains and outs[-1] == b""
    assert _second(emu) - before == 0


@pytest.mark.parametrize("damage", [1, 2, 5])
def test_xhuff_indexed_decode_survives_a_wrong_index(emu, oracle, ms, damage):
    plains = [ms.generate("text", 70000, 3), ms.generate("runs", 40000, 4), b"abc" * 3000]
    comps = [oracle.compress("xpress_huff", p) for p in plains]
    before = _second(emu)
    st, ol, outs = emu_batch(emu, "emu_xhuff_decode_indexed", comps, [len(p) for p in plains], extra=(ctypes.c_uint32(damage), ctypes.c_uint32(1)))
    assert not st.any() and outs == plains
    assert _second(emu) - before >= 1


def test_xhuff_indexed_status_equals_oracle_on_bad_input(emu, oracle, ms):
    rnd = random.Random(6)
    cases = []
    for kind in ("text", "runs"):
        n = 20000
        plain = ms.generate(kind, n, 9)
        comp = oracle.compress("xpress_huff", plain)
        cases += [(comp, cap) for cap in (n, n + 100, n - 1, n // 2, 0)]
        cases += [(comp[:-t], n) for t in (1, 2, 3, len(comp) // 2)]
        cases += [(comp + b"\0\0", n), (comp + b"\0\0", n + 10), (comp + b"\x55", n)]
        for _ in range(40):
            b = bytearray(comp)
            for _ in range(rnd.randint(1, 3)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            cases.append((bytes(b), n + rnd.choice((0, 64))))
    st, ol, outs = emu_batch(emu, "emu_xhuff_decode_indexed", [c for c, _ in cases], [cap for _, cap in cases], extra=(ctypes.c_uint32(0), ctypes.c_uint32(0)))
    for i, (src, cap) in enumerate(cases):
        rc, exp = oracle.decompress("xpress_huff", src, cap)
        assert st[i] == rc, (i, st[i], rc)
        if rc == 0:
            assert outs[i] == exp


def test_xhuff_encoder_writes_the_index_the_walk_finds(emu, oracle, ms):
    """mscomp_gpu_xpress_huff_compress_batch_indexed: the encoder's own restart index equals, entry for entry, the one the serial
    walk finds in the stream it wrote (multi-chunk buffers, the literal-only fallback, long matches with raw length bytes, tiny
    buffers), and the streams are the ones the plain call writes."""
    plains = corpus(ms) + [ms.generate("text", 200000, 5), ms.generate("random", 70000, 6), ms.generate("runs", 140000, 7), b"a", b"ab" * 40000]
    caps = [len(p) + 34 + 258 + 258 * (len(p) // 65536) + 16 for p in plains]
    st, ol, outs = emu_batch(emu, "emu_xhuff_encode_index_check", plains, caps)
    assert not st.any() and emu_batch.last_return == 0
    st2, ol2, outs2 = emu_batch(emu, "emu_xhuff_encode", plains, caps)
    assert outs == outs2
    for p, c in zip(plains, outs):
        rc, back = oracle.decompress("xpress_huff", c, len(p))
        assert rc == 0 and back == p


def _emu_stream_decode(emu, comp, cap):
    import numpy as np
    a, addr = padded(comp)
    out = np.zeros(cap + 2 * PAD, dtype=np.uint8)
    total, info = ctypes.c_uint64(0), ctypes.c_uint32(0)
    emu.emu_xhuff_stream_decode.restype = ctypes.c_uint32
    handled = emu.emu_xhuff_stream_decode(ctypes.c_void_p(addr), ctypes.c_uint32(len(comp)), ctypes.c_void_p(out.ctypes.data + PAD),
                                          ctypes.c_uint64(cap), ctypes.byref(total), ctypes.byref(info))
    return handled, out[PAD:PAD + total.value].tobytes(), info.value


def test_xhuff_chunk_parallel_stream_decode(emu, oracle, ms):
    """One multi-chunk stream: chunk starts recognised by their complete code-length tables, one lane per chunk, chain stitched from
    offset 0.  Reference-made streams have matches that reach into the previous chunk (executed in order); streams made by our
    encoder do not (all chunks at once).  Streams the path cannot vouch for must be declined, never decoded wrongly."""
    for kind, n in (("text", 200000), ("text", 65536 * 3), ("runs", 150000), ("random", 140000), ("text", 65537), ("runs", 400000)):
        plain = ms.generate(kind, n, 7 + n % 5)
        comp = oracle.compress("xpress_huff", plain)
        for in_order, lane_parse in ((1, 1), (0, 1), (0, 2), (1, 0), (0, 0)):  # the in-order block executor / pointer jumping (lz_jump.cuh, the default);
            emu.emu_xhuff_stream_in_order(in_order)                  # the batch parser one lane per chunk (1) / a warp per chunk, lane 0 walks (2) / speculative rounds (0, the default)
            emu.emu_xhuff_stream_lane_parse(lane_parse)
            handled, got, info = _emu_stream_decode(emu, comp, n)
            assert handled == 1 and got == plain, (kind, n, handled, info, in_order, lane_parse)
            assert (info >> 16) & 0x7FFF == (n + 65535) // 65536
        # our own encoder: window reset per chunk -> independent chunks
        st, ol, outs = emu_batch(emu, "emu_xhuff_encode", [plain], [len(plain) + 34 + 258 + 258 * (len(plain) // 65536) + 16])
        assert not st.any()
        for own_warps in (1, 0):                                     # every chunk on its own warp (many chunks) / pointer jumping (few chunks)
            emu.emu_xhuff_stream_in_order(own_warps)
            handled, got, info = _emu_stream_decode(emu, outs[0], n)
            assert handled == 1 and got == plain and (info >> 31) == 1, (kind, n, info, own_warps)
        # too little room, a damaged table, a truncated stream: declined (the serial decoder reports those)
        assert _emu_stream_decode(emu, comp, n - 1)[0] == 0
        bad = bytearray(comp); bad[3] ^= 0x10
        assert _emu_stream_decode(emu, bytes(bad), n)[0] == 0
        h, g, _ = _emu_stream_decode(emu, comp[:len(comp) - 5], n)
        assert h == 0 or g == oracle.decompress("xpress_huff", comp[:len(comp) - 5], n)[1]
    # very long matches across chunk starts (rounds longer than the ring's span run their tokens in order)
    taken = 0
    for plain in (bytes(200000), b"ab" * 100000, ms.generate("text", 3000, 2) * 70, bytes(range(256)) * 1000, ms.generate("runs", 400000, 3)):
        comp = oracle.compress("xpress_huff", plain)
        handled, got, info = _emu_stream_decode(emu, comp, len(plain))
        assert handled == 0 or got == plain, (len(plain), info)      # (a chunk that uses ONE symbol has an incomplete code: declined)
        taken += handled
    assert taken >= 2
    # the speculative rounds carry the load: on text nearly every symbol is taken by a round, the serial routine only sees length escapes and chunk ends
    cnt = (ctypes.c_uint64 * 2)()
    emu.emu_xhuff_chunk_warp_counts(cnt)
    plain = ms.generate("text", 200000, 3)
    assert _emu_stream_decode(emu, oracle.compress("xpress_huff", plain), len(plain))[1] == plain
    emu.emu_xhuff_chunk_warp_counts(cnt)
    assert cnt[0] > 20 * cnt[1] and cnt[0] > 30000, list(cnt)
    # a single chunk takes the same path (one warp walks it)
    for kind, n in (("text", 50000), ("text", 65536), ("runs", 30000), ("random", 3000), ("text", 700)):
        plain = ms.generate(kind, n, 1)
        comp = oracle.compress("xpress_huff", plain)
        handled, got, info = _emu_stream_decode(emu, comp, n)
        assert handled == 1 and got == plain, (kind, n, info)
        assert _emu_stream_decode(emu, comp, n - 1)[0] == 0
        for t in (1, 2, 7):
            h, g, _ = _emu_stream_decode(emu, comp[:len(comp) - t], n)
            assert h == 0 or (0, g) == oracle.decompress("xpress_huff", comp[:len(comp) - t], n), (kind, n, t)


def _emu_xpress_stream_decode(emu, comp, cap):
    a, addr = padded(comp)
    out = np.zeros(cap + 2 * PAD, dtype=np.uint8)
    total = ctypes.c_uint64(0)
    info = (ctypes.c_uint32 * 4)()
    emu.emu_xpress_stream_decode.restype = ctypes.c_uint32
    handled = emu.emu_xpress_stream_decode(ctypes.c_void_p(addr), ctypes.c_uint32(len(comp)), ctypes.c_void_p(out.ctypes.data + PAD),
                                           ctypes.c_uint64(cap), ctypes.byref(total), info)
    return handled, out[PAD:PAD + total.value].tobytes(), list(info)


def test_xpress_speculative_stream_decode(emu, oracle, ms):
    """One plain Xpress stream, flag groups found by speculation (kernels/xpress_spec.cuh): the group that would start at EVERY offset in
    every half-byte state, the true chain linked up block by block, a restart index from it, the indexed parser and the order-free
    token execution.  Reference-made streams of all kinds decode bit-exactly; streams the path cannot vouch for are declined
    (the serial decoder reports those), never decoded wrongly."""
    rnd = random.Random(11)
    plains = [ms.generate(kind, n, 3 + n % 7) for kind, n in (("text", 200000), ("runs", 150000), ("random", 70000), ("text", 65537), ("runs", 400000),
                                                              ("text", 9000), ("text", 300))]
    plains += [bytes(200000), b"ab" * 100000, ms.generate("text", 3000, 2) * 70, bytes(range(256)) * 1000,
               bytes(rnd.randrange(4) for _ in range(120000)),                     # many long matches: the shared length nibbles are busy
               b"".join(bytes([rnd.randrange(256)]) * rnd.choice((1, 2, 9, 10, 11, 24, 25, 26, 300, 70000)) for _ in range(3000))]
    for plain in plains:
        n = len(plain)
        comp = oracle.compress("xpress", plain)
        handled, got, info = _emu_xpress_stream_decode(emu, comp, n)
        if len(comp) < 64:
            assert handled == 0
            continue
        assert handled == 1 and got == plain, (n, len(comp), handled, info)
        assert info[2] == (info[1] + 15) // 16 and info[0] == (len(comp) - 3 + 8191) // 8192
        assert _emu_xpress_stream_decode(emu, comp, n + 1000)[1] == plain         # more room than needed
        assert _emu_xpress_stream_decode(emu, comp, n - 1)[0] == 0                 # too little: declined
        # damaged streams: declined, or exactly what the oracle makes of them
        for t in (1, 2, 5, len(comp) // 2):
            cut = comp[:len(comp) - t]
            h, g, _ = _emu_xpress_stream_decode(emu, cut, n)
            assert h == 0 or (0, g) == oracle.decompress("xpress", cut, n), (n, t)
        for _ in range(6):
            b = bytearray(comp)
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            h, g, _ = _emu_xpress_stream_decode(emu, bytes(b), n)
            assert h == 0 or (0, g) == oracle.decompress("xpress", bytes(b), n), n
    # streams of this library's own encoders decode the same way
    for name, extra in (("emu_xpress_encode", ()), ("emu_xpress_encode_exact", (ctypes.c_uint32(16384), ctypes.c_void_p(0), ctypes.c_uint32(0)))):
        plain = ms.generate("text", 150000, 5)
        st, ol, outs = emu_batch(emu, name, [plain], [MAXSZ["xpress"](len(plain))], extra=extra)
        assert not st.any()
        handled, got, _ = _emu_xpress_stream_decode(emu, outs[0], len(plain))
        assert handled == 1 and got == plain


def test_xhuff_chunk_warp_parser_fuzz(emu, oracle, ms):
    """The warp-per-chunk parser (xhuff_chunk_warp.cuh: speculation on the start bit, reader state from the bits consumed) against the oracle
    on random inputs of several kinds and sizes, whole and damaged: a handled stream must be exactly what the oracle decodes."""
    rnd = random.Random(2024)
    emu.emu_xhuff_stream_in_order(0)
    emu.emu_xhuff_stream_lane_parse(0)
    taken = 0
    for it in range(24):
        kind = rnd.choice(["text", "runs", "random", "sym4", "mix", "rep"])
        n = rnd.choice([700, 3000, 20000, 65536, 70000, 131072])
        if kind in ("text", "runs", "random"):
            plain = ms.generate(kind, n, rnd.randrange(1000))
        elif kind == "sym4":
            plain = bytes(rnd.randrange(4) for _ in range(n))
        elif kind == "rep":
            plain = (ms.generate("text", rnd.choice([5, 37, 300, 3000]), rnd.randrange(99)) * (n // 5 + 1))[:n]
        else:
            plain = b"".join(rnd.choice([ms.generate("text", 3000, rnd.randrange(99)), bytes(rnd.randrange(2000)),
                                         bytes(rnd.getrandbits(8) for _ in range(500)), b"ab" * rnd.randrange(400)]) for _ in range(n // 2000 + 1))[:n]
        n = len(plain)
        comp = oracle.compress("xpress_huff", plain)
        handled, got, info = _emu_stream_decode(emu, comp, n)
        assert handled == 0 or got == plain, (kind, n, info)
        taken += handled
        for _ in range(2):
            b = bytearray(comp)
            if rnd.random() < 0.5:
                b = b[:len(b) - rnd.randrange(1, 9)]
            else:
                for _ in range(rnd.randint(1, 3)):
                    b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            cap = n + rnd.choice((0, 0, 64))
            h2, g2, _ = _emu_stream_decode(emu, bytes(b), cap)
            assert h2 == 0 or (0, g2) == oracle.decompress("xpress_huff", bytes(b), cap), (kind, n)
    assert taken >= 18
