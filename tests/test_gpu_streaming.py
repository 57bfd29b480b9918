"""Streaming API (lznt1_deflate / lznt1_inflate / ms_* dispatch, MSCOMP_NONE copy streams) against the compiled reference,
call by call: both libraries are driven with the same random sequence of input pieces, output capacities and flush flags and
must return the same status, consume and produce the same number of bytes in EVERY call, and produce the same bytes."""
import ctypes
import random

import pytest

pytestmark = pytest.mark.gpu

from stream_util import FINISH, FLUSH, NO_FLUSH, POSSIBLE_END, STREAM_END, Driver, Stream, drive_pair  # noqa: E402,F401


def drive(ms, reflib, prefix, fmt, deflate, data, rnd, max_piece, max_cap, use_flush):
    f = fmt if prefix == "ms_" else None
    return drive_pair(Driver(ms.lib(), prefix, deflate, f), Driver(reflib.lib, prefix, deflate, f), deflate, data, rnd, max_piece, max_cap, use_flush)


@pytest.mark.parametrize("prefix", ["lznt1_", "ms_"])
def test_deflate_matches_reference_call_by_call(ms, reflib, oracle, prefix):
    rnd = random.Random(5)
    for kind, n in (("text", 200000), ("random", 50000), ("runs", 120000), ("text", 4096 * 7), ("text", 5), ("text", 0)):
        plain = ms.generate(kind, n, 3 + n % 5)
        for max_piece, max_cap, use_flush in ((70000, 90000, False), (5000, 3000, True), (300, 9000, True), (20000, 100, False)):
            comp, rc = drive(ms, reflib, prefix, 2, True, plain, rnd, max_piece, max_cap, use_flush)
            assert rc == STREAM_END
            if not use_flush:
                assert comp == oracle.compress("lznt1", plain)
            assert oracle.decompress("lznt1", comp, n) == (0, plain)


@pytest.mark.parametrize("prefix", ["lznt1_", "ms_"])
def test_inflate_matches_reference_call_by_call(ms, reflib, oracle, prefix):
    rnd = random.Random(6)
    a, b, c = ms.generate("text", 100, 1), ms.generate("text", 50, 2), ms.generate("text", 90000, 3)
    flushed = b"".join(oracle.compress("lznt1", x) for x in (a, b, c))                  # short chunks in mid-stream
    streams = [(oracle.compress("lznt1", ms.generate(kind, n, 9 + n % 7)), ms.generate(kind, n, 9 + n % 7))
               for kind, n in (("text", 300000), ("random", 40000), ("runs", 150000), ("text", 4096), ("text", 3))]
    streams.append((flushed, a + b + c))
    for comp, plain in streams:
        variants = [comp, comp + b"\0\0", comp + b"\0", comp[:-3], comp + b"\0\0\x01"]
        bad = bytearray(comp)
        if len(bad) > 100:
            bad[len(bad) // 2] ^= 0x10
            variants.append(bytes(bad))
            hdr = bytearray(comp); hdr[1] ^= 0x40
            variants.append(bytes(hdr))
        for v in variants:
            for max_piece, max_cap in ((70000, 90000), (5000, 3000), (300, 5000), (20000, 100), (1, 4096)):
                if max_piece == 1 and len(v) > 6000:
                    continue                                  # one byte per call: small streams only
                out, rc = drive(ms, reflib, prefix, 2, False, v, rnd, max_piece, max_cap, False)
                if v is comp:
                    assert out == plain and rc >= 0


def test_copy_streams_and_unsupported_formats(ms, reflib):
    rnd = random.Random(7)
    data = ms.generate("text", 30000, 1)
    for deflate in (True, False):
        out, rc = drive(ms, reflib, "ms_", 0, deflate, data, rnd, 4000, 3000, False) if deflate else (None, None)
    # formats without a stream implementation here: an error from init, never a crash
    s = Stream()
    for fmt in (1, 3, 4, 5, 77):
        f = ms.lib().ms_deflate_init; f.restype = ctypes.c_int
        g = ms.lib().ms_inflate_init; g.restype = ctypes.c_int
        assert f(ctypes.c_int(fmt), ctypes.byref(s)) < 0 and g(ctypes.c_int(fmt), ctypes.byref(s)) < 0
