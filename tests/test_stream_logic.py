"""The product's LZNT1 stream state machine (ms-compress_b200/csrc/stream_logic.h) on the CPU: instantiated with the oracle's chunk
codec (tests/emu/stream_host.cpp) and driven against the compiled reference call by call -- same random schedule of input pieces,
output capacities and flush flags, same status / bytes consumed / bytes produced in every call.  The GPU suite runs the same
schedules through the product library itself (tests/test_gpu_streaming.py)."""
import ctypes
import os
import random
import subprocess

import pytest
from conftest import ROOT
from stream_util import STREAM_END, Driver, Stream, drive_pair


@pytest.fixture(scope="module")
def host():
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d, "streamhost"], check=True)
    return ctypes.CDLL(os.path.join(d, "libstreamhost.so"))


def drive(host, reflib, deflate, data, rnd, max_piece, max_cap, use_flush):
    return drive_pair(Driver(host, "hs_lznt1_", deflate), Driver(reflib.lib, "lznt1_", deflate), deflate, data, rnd, max_piece, max_cap, use_flush)


def test_deflate_state_machine(host, reflib, oracle, ms):
    rnd = random.Random(5)
    for kind, n in (("text", 120000), ("random", 30000), ("runs", 90000), ("text", 4096 * 5), ("text", 5), ("text", 0)):
        plain = ms.generate(kind, n, 3 + n % 5)
        for max_piece, max_cap, use_flush in ((70000, 90000, False), (5000, 3000, True), (300, 9000, True), (20000, 100, False), (7, 5000, True)):
            comp, rc = drive(host, reflib, True, plain, rnd, max_piece, max_cap, use_flush)
            assert rc == STREAM_END
            if not use_flush:
                assert comp == oracle.compress("lznt1", plain)
            assert oracle.decompress("lznt1", comp, n) == (0, plain)


def test_inflate_state_machine(host, reflib, oracle, ms):
    rnd = random.Random(6)
    a, b, c = ms.generate("text", 100, 1), ms.generate("text", 50, 2), ms.generate("text", 60000, 3)
    flushed = b"".join(oracle.compress("lznt1", x) for x in (a, b, c))                  # short chunks in mid-stream
    streams = [(oracle.compress("lznt1", ms.generate(kind, n, 9 + n % 7)), ms.generate(kind, n, 9 + n % 7))
               for kind, n in (("text", 150000), ("random", 30000), ("runs", 100000), ("text", 4096), ("text", 3))]
    streams.append((flushed, a + b + c))
    for comp, plain in streams:
        variants = [comp, comp + b"\0\0", comp + b"\0", comp[:-3], comp + b"\0\0\x01", comp + b"\x07"]
        if len(comp) > 100:
            bad = bytearray(comp); bad[len(bad) // 2] ^= 0x10
            hdr = bytearray(comp); hdr[1] ^= 0x40
            variants += [bytes(bad), bytes(hdr)]
        for v in variants:
            for max_piece, max_cap in ((70000, 90000), (5000, 3000), (300, 5000), (20000, 100), (1, 4096)):
                if max_piece == 1 and len(v) > 6000:
                    continue
                out, rc = drive(host, reflib, False, v, rnd, max_piece, max_cap, False)
                if v is comp:
                    assert out == plain and rc >= 0


def test_copy_streams(host, reflib, ms):
    rnd = random.Random(7)
    data = ms.generate("text", 30000, 1)
    # the copy engine shares one init / end for both directions (src/mscomp.cpp:34-61): drive it through small wrappers
    class Copy(Driver):
        def __init__(self, lib, deflate, ours):
            self.lib, self.deflate = lib, deflate
            self.s = Stream()
            if ours:
                init, self.f, self.end = lib.hs_copy_init, (lib.hs_copy_deflate if deflate else lib.hs_copy_inflate), lib.hs_copy_end
                self.rc_init = init(ctypes.byref(self.s))
            else:
                init = lib.ms_deflate_init if deflate else lib.ms_inflate_init
                self.f = lib.ms_deflate if deflate else lib.ms_inflate
                self.end = lib.ms_deflate_end if deflate else lib.ms_inflate_end
                self.rc_init = init(ctypes.c_int(0), ctypes.byref(self.s))
            for fn in (self.f, self.end):
                fn.restype = ctypes.c_int
            self.out = bytearray()
            self.pending = b""
    for deflate in (True, False):
        out, rc = drive_pair(Copy(host, deflate, True), Copy(reflib.lib, deflate, False), deflate, data, rnd, 4000, 3000, False)
        assert out == data


def test_one_shot_decompress_host_logic(host, reflib, oracle, ms):
    """Header walk + status replay of the one-shot lznt1_decompress (csrc/lznt1_host.h) over the oracle's chunk decoder: same status
    and bytes as the compiled reference on valid, truncated, padded and damaged streams and on every kind of capacity."""
    rnd = random.Random(8)
    f = host.hs_lznt1_decompress
    f.restype = ctypes.c_int

    def ours(src, cap):
        out = ctypes.create_string_buffer(max(cap, 1))
        n = ctypes.c_size_t(cap)
        rc = f(src, ctypes.c_size_t(len(src)), out, ctypes.byref(n))
        return rc, (out.raw[:n.value] if rc == 0 else None)

    a, b, c = ms.generate("text", 100, 1), ms.generate("text", 50, 2), ms.generate("text", 9000, 3)
    flushed = b"".join(oracle.compress("lznt1", x) for x in (a, b, c))
    checked = 0
    for plain in [ms.generate(k, n, 5 + n) for k in ("text", "runs", "random") for n in (0, 1, 3, 4095, 4096, 4097, 30000)] + [a + b + c]:
        comp = flushed if plain == a + b + c else oracle.compress("lznt1", plain)
        n = len(plain)
        cases = [(comp, cap) for cap in (n, n + 1, n + 100, max(0, n - 1), n // 2, 0, 4 * n + 7)]
        cases += [(comp + t, n + d) for t in (b"\0", b"\x55", b"\0\0", b"\0\0\x01", b"\0" * 5) for d in (0, 10)]
        cases += [(comp[:-t], n) for t in (1, 2, 3, max(1, len(comp) // 2)) if t <= len(comp)]
        for _ in range(25):
            d = bytearray(comp)
            for _ in range(rnd.randint(1, 3)):
                if d:
                    d[rnd.randrange(len(d))] ^= 1 << rnd.randrange(8)
            cases.append((bytes(d), n + rnd.choice((0, 64))))
        for src, cap in cases:
            assert ours(src, cap) == reflib.call("lznt1_decompress", src, cap), (len(src), cap)
            checked += 1
    assert checked > 500


def test_xpress_inflate_state_machine(host, reflib, oracle, ms):
    """xpress_inflate (kernels/xpress_stream.cuh + the host logic of stream_logic.h) against the compiled reference, call by call:
    whole streams, streams with trailing bytes, truncated and damaged streams; pieces down to one byte, output room down to zero."""
    rnd = random.Random(8)
    streams = [(oracle.compress("xpress", ms.generate(kind, n, 9 + n % 7)), ms.generate(kind, n, 9 + n % 7))
               for kind, n in (("text", 150000), ("random", 20000), ("runs", 100000), ("text", 4096), ("text", 3), ("runs", 9000))]
    streams.append((oracle.compress("xpress", b"ab" * 50000), b"ab" * 50000))           # long lengths: u16 extensions
    streams.append((oracle.compress("xpress", bytes(700000)), bytes(700000)))           # u32 length extension
    for si, (comp, plain) in enumerate(streams):
        variants = [comp, comp + b"\0\0", comp[:-3], comp[:len(comp) // 2], comp + b"\xff\xff\xff\xff", comp + b"\x07"]
        if len(comp) > 100:
            for _ in range(3):
                bad = bytearray(comp); bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
                variants.append(bytes(bad))
        short_matches = si in (0, 3, 4)             # text: with the slip reproduced, a match longer than every later output room never ends
        for v in variants:
            for max_piece, max_cap in ((70000, 90000), (5000, 3000), (300, 5000), (20000, 100), (7, 50), (1, 4096)):
                if max_piece <= 7 and len(v) > 6000:
                    continue
                if not short_matches:
                    # long matches: only calls with room for everything can be compared with the reference (see above)
                    a, b = Driver(host, "hs_xpressq_", False), Driver(reflib.lib, "xpress_", False)
                    pos = 0
                    while True:
                        piece = v[pos:pos + rnd.randint(0, max_piece)]; pos += len(piece)
                        ra, rb = a.call(piece, len(plain) + 64, 0), b.call(piece, len(plain) + 64, 0)
                        assert ra == rb
                        if ra[0] < 0 or (pos >= len(v) and not piece):
                            break
                    assert bytes(a.out) == bytes(b.out) and a.finish() == b.finish()
                    if max_cap < 2000:
                        continue
                    out, rc = bytes(a.out), ra[0]
                else:
                    out, rc = drive_pair(Driver(host, "hs_xpressq_", False), Driver(reflib.lib, "xpress_", False), False, v, rnd, max_piece, max_cap, False)
                # (hs_xpressq_: with the reference's slip in the owed-match bookkeeping reproduced -- src/xpress_decompress.cpp:261-263 --
                # so that schedules where a match is owed over three or more calls can be compared too; there the reference's
                # OUTPUT is wrong, which the second run below, without the slip, shows by getting the right bytes)
                if v is comp:
                    assert rc >= 0
                    d = Driver(host, "hs_xpress_", False)
                    pos = 0
                    for _ in range(200000):
                        piece = v[pos:pos + rnd.randint(0, max_piece)]; pos += len(piece)
                        r = d.call(piece, rnd.choice((0, 1, 2, 3, rnd.randint(0, max_cap), max_cap)), 0)
                        assert r[0] >= 0
                        if pos >= len(v) and not d.pending and len(d.out) >= len(plain):
                            break
                    assert bytes(d.out) == plain


def test_xpress_deflate_is_as_unfinished_as_upstream(host, reflib):
    for lib, stem in ((host, "hs_xpress_"), (reflib.lib, "xpress_")):
        s = Stream()
        f = getattr(lib, stem + "deflate_init"); f.restype = ctypes.c_int
        g = getattr(lib, stem + "deflate"); g.restype = ctypes.c_int
        h = getattr(lib, stem + "deflate_end"); h.restype = ctypes.c_int
        assert f(ctypes.byref(s)) == -4 and g(ctypes.byref(s), ctypes.c_int(0)) == -2 and h(ctypes.byref(s)) == -2
