"""GPU chunk index of an LZNT1 stream (kernels/lznt1_index.cuh) under the SIMT emulator, against the serial header walk
it replaces (the reference's driver, src/lznt1_decompress.cpp:254-273 with the header read of :127-135)."""
import ctypes
import random

import numpy as np
import pytest
from emu_util import P, padded

CLEAN, ZERO, GARBAGE, TERM_OK, TERM_BAD, TRUNC, BADSIG = range(7)


def host_walk(c):
    """The serial walk: chunk offsets, how the stream ended, bytes consumed."""
    offs, ip, n = [], 0, len(c)
    while n - ip >= 2:
        h = c[ip] | (c[ip + 1] << 8)
        if h == 0:
            return offs, (TERM_OK if n - ip == 2 else TERM_BAD), ip
        sz = (h & 0xFFF) + 3
        if sz > n - ip:
            return offs, TRUNC, ip
        if (h & 0x7000) != 0x3000:
            return offs, BADSIG, ip
        offs.append(ip)
        ip += sz
    return offs, (CLEAN if n - ip == 0 else (ZERO if c[ip] == 0 else GARBAGE)), ip


def gpu_index(emu, c, S, range_segs):
    a, addr = padded(c)
    cap = len(c) // 3 + 2
    offs = np.zeros(cap + 1, dtype=np.uint64)
    st = np.zeros(4, dtype=np.uint64)
    emu.emu_lznt1_index.restype = ctypes.c_int
    rc = emu.emu_lznt1_index(ctypes.c_void_p(addr), ctypes.c_uint64(len(c)), ctypes.c_uint32(S), ctypes.c_uint32(range_segs), P(offs),
                             ctypes.c_uint64(cap), P(st))
    assert rc == 0
    done, tail, ip, n = (int(x) for x in st)
    assert done == 1
    return [int(x) for x in offs[:n]], tail, ip


def check(emu, c, S=4112, range_segs=5):
    assert gpu_index(emu, c, S, range_segs) == host_walk(c), (len(c), S, range_segs)


def synth_stream(rnd, nchunks, lo, hi, stored_ratio=0.2):
    """Headers with plausible signatures over random payload (the index does not look at payloads)."""
    out = bytearray()
    for _ in range(nchunks):
        sz = rnd.randint(lo, hi)
        h = (0x3000 if rnd.random() < stored_ratio else 0xB000) | (sz - 3)
        out += bytes([h & 255, h >> 8]) + bytes(rnd.getrandbits(8) for _ in range(sz - 2))
    return bytes(out)


def test_index_of_real_streams(emu, oracle, ms):
    for kind in ("text", "random", "runs"):
        for n in (0, 1, 5000, 70000, 300000):
            c = oracle.compress("lznt1", ms.generate(kind, n, 7 + n))
            for S, rs in ((4112, 3), (8192, 1000), (65536, 2)):
                check(emu, c, S, rs)


def test_index_of_synthetic_streams(emu):
    rnd = random.Random(5)
    for lo, hi, n in ((3, 3, 5000), (3, 40, 3000), (2000, 2700, 60), (4098, 4098, 30), (3, 4098, 200)):
        c = synth_stream(rnd, n, lo, hi)
        for S, rs in ((4112, 4), (4112, 1), (16384, 3)):
            check(emu, c, S, rs)
            check(emu, c + b"\x00", S, rs)                 # one stray zero byte
            check(emu, c + b"\x07", S, rs)                 # one stray non-zero byte
            check(emu, c + b"\x00\x00", S, rs)             # terminator
            check(emu, c + b"\x00\x00\x01", S, rs)         # terminator followed by data
            check(emu, c[:-1], S, rs)                      # truncated last chunk
    for n in range(0, 9):                                  # tiny inputs, every prefix of a valid 1-chunk stream + zeros
        check(emu, bytes([3, 0x30, 1, 2, 3, 4, 0, 0, 0])[:n])
        check(emu, bytes(n))


def test_index_stops_like_the_walk_on_damaged_headers(emu, oracle, ms):
    rnd = random.Random(9)
    base = oracle.compress("lznt1", ms.generate("text", 150000, 3))
    offs, _, _ = host_walk(base)
    for _ in range(12):
        c = bytearray(base)
        at = rnd.choice(offs)
        mode = rnd.randrange(4)
        if mode == 0:
            c[at + 1] ^= 0x40                              # signature bits
        elif mode == 1:
            c[at] = c[at + 1] = 0                          # terminator in mid-stream
        elif mode == 2:
            c[at] ^= rnd.randint(1, 255)                   # size: the walk goes astray from here
        else:
            c[at + 1] = (c[at + 1] & 0xF0) | rnd.randrange(16)
        check(emu, bytes(c), 4112, rnd.choice((1, 2, 7)))
        check(emu, bytes(c), 8192, 1000)


@pytest.mark.parametrize("S", [4112, 65536])
def test_stream_ending_on_a_segment_boundary(emu, S):
    rnd = random.Random(1)
    c = synth_stream(rnd, 40, 1000, 3000)
    pad = (-len(c)) % S
    while pad < 3:
        pad += S
    h = 0x3000 | (pad - 3)
    c += bytes([h & 255, h >> 8]) + bytes(pad - 2)
    assert len(c) % S == 0
    check(emu, c, S, 2)
    check(emu, c + b"\x00", S, 2)
    check(emu, c[:-1], S, 2)
