"""The LZNT1 CUDA kernels, compiled for the CPU through tests/emu/simt_emu.h, against the oracle.
(The same kernels run on the B200 in tests/test_gpu_lznt1.py; this file keeps their logic covered without a GPU.)"""
import ctypes
import random

import pytest
from emu_util import emu_lznt1_decode, emu_lznt1_encode, lznt1_walk

SIZES = [1, 2, 3, 4, 7, 8, 9, 16, 17, 31, 32, 33, 100, 4095, 4096, 4097, 8193, 65536, 150000]


KERNELS = ["emu_lznt1_decode"]      # warp per chunk


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("kind", ["text", "runs", "random"])
def test_decode_matches_oracle(emu, oracle, ms, kind, kernel):
    for n in SIZES:
        plain = ms.generate(kind, n, 42 + n)
        comp = oracle.compress("lznt1", plain)
        st, sz, chunks = emu_lznt1_decode(emu, comp, kernel=kernel)
        assert not st.any() and b"".join(chunks) == plain, (kind, n)


@pytest.mark.parametrize("heavy_work", [32768, 0], ids=["default", "all-chunks-lazy"])
@pytest.mark.parametrize("kind", ["text", "runs", "random"])
def test_encode_is_byte_identical_to_oracle(emu, oracle, ms, kind, heavy_work):
    """heavy_work=0 sends every chunk through the lazy (search-on-visit) parse that heavy chunks take."""
    for n in SIZES:
        plain = ms.generate(kind, n, 42 + n)
        got, _ = emu_lznt1_encode(emu, plain, heavy_work=heavy_work)
        assert got == oracle.compress("lznt1", plain), (kind, n)


def test_encode_heavy_chunks(emu, oracle):
    """Tiny alphabets and interleaved runs: huge hash buckets, the default threshold routes them to the lazy parse."""
    rnd = random.Random(11)
    cases = [bytes(rnd.choice(b"ACGT") for _ in range(3 * 4096 + 100)),
             b"".join(bytes([rnd.choice(b"\x00\x00\x00\xff\x01")]) * rnd.randint(1, 40) for _ in range(700)),
             b"".join((b"\x00" * rnd.randint(3, 300)) + bytes([rnd.randint(1, 255)]) for _ in range(120))]
    for plain in cases:
        for hw in (32768, 0, 1 << 30):
            got, _ = emu_lznt1_encode(emu, plain, heavy_work=hw)
            assert got == oracle.compress("lznt1", plain)


def test_encode_degenerate_inputs(emu, oracle):
    for plain in (bytes(20000), b"ab" * 5000, b"abc" * 100, bytes(range(97, 123)), bytes(range(256)) * 40, b"a" * 4096 + b"b" * 4096):
        got, _ = emu_lznt1_encode(emu, plain)
        assert got == oracle.compress("lznt1", plain)


@pytest.mark.parametrize("kernel", KERNELS)
def test_decode_status_on_corrupted_chunks(emu, oracle, ms, kernel):
    """Bit flips: per-chunk status and output equal the oracle's chunk decoder (src/lznt1_decompress.cpp:37-121)."""
    rnd = random.Random(3)
    f = oracle.lib.orc_lznt1_decompress_chunk
    f.restype = ctypes.c_int
    for kind in ("text", "runs"):
        plain = ms.generate(kind, 40000, 5)
        comp = oracle.compress("lznt1", plain)
        offs = lznt1_walk(comp)
        for _ in range(60):
            b = bytearray(comp)
            for _ in range(rnd.randint(1, 4)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
            b = bytes(b)
            st, sz, chunks = emu_lznt1_decode(emu, b, offs, kernel=kernel)
            for ci, p in enumerate(offs):
                h = b[p] | (b[p + 1] << 8)
                n = (h & 0xFFF) + 1
                if h == 0 or (h & 0x7000) != 0x3000:
                    exp, eo = -3, None
                elif p + 2 + n > len(b):
                    exp, eo = -5, None
                elif not h & 0x8000:
                    exp, eo = 0, b[p + 2:p + 2 + n]
                else:
                    o = ctypes.create_string_buffer(4096)
                    ol = ctypes.c_size_t(0)
                    exp = f(b[p + 2:p + 2 + n], ctypes.c_size_t(n), o, ctypes.c_size_t(4096), ctypes.byref(ol))
                    eo = o.raw[:ol.value] if exp == 0 else None
                assert st[ci] == exp
                if exp == 0:
                    assert chunks[ci] == eo
