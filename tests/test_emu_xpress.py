"""Xpress / Xpress-Huffman CUDA kernels compiled for the CPU (tests/emu) against the oracle.
Decoders: bit-exact output and identical status.  Encoders: the stream must decode with the oracle (= the reference's
decoder restated) back to the input, stay inside the reference's max_compressed_size bound and near the reference's ratio."""
import random

import pytest
from conftest import MAXSZ
from emu_util import emu_batch

KERNELS = {"xpress_huff": ("emu_xhuff_decode", "emu_xhuff_encode"), "xpress": ("emu_xpress_decode", "emu_xpress_encode")}
SIZES = [1, 2, 3, 4, 5, 31, 32, 33, 100, 5000, 65535, 65536, 65537, 140000]


def corpus(ms):
    out = []
    for kind in ("text", "runs", "random"):
        out += [ms.generate(kind, n, 42 + n) for n in SIZES]
    out += [b"abc" * 100, bytes(100000), bytes(range(97, 123)), b"ab" * 40000, bytes(range(256)) * 300]
    return out


DECODERS = {"xpress_huff": ["emu_xhuff_decode", "emu_xhuff_decode2"], "xpress": ["emu_xpress_decode", "emu_xpress_decode2"]}   # fused, parse+apply


@pytest.mark.parametrize("variant", [0, 1])
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


@pytest.mark.parametrize("variant", [0, 1])
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
