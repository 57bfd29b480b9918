"""Streaming API (lznt1_deflate / lznt1_inflate / ms_* dispatch, MSCOMP_NONE copy streams) against the compiled reference,
call by call: both libraries are driven with the same random sequence of input pieces, output capacities and flush flags and
must return the same status, consume and produce the same number of bytes in EVERY call, and produce the same bytes."""
import ctypes
import random

import pytest

pytestmark = pytest.mark.gpu

NO_FLUSH, FLUSH, FINISH = 0, 2, 4
STREAM_END, POSSIBLE_END = 1, 2


class Stream(ctypes.Structure):
    _fields_ = [("format", ctypes.c_int), ("compressing", ctypes.c_bool), ("in_", ctypes.c_void_p), ("in_avail", ctypes.c_size_t),
                ("in_total", ctypes.c_size_t), ("out", ctypes.c_void_p), ("out_avail", ctypes.c_size_t), ("out_total", ctypes.c_size_t),
                ("error", ctypes.c_char * 256), ("warning", ctypes.c_char * 256), ("state", ctypes.c_void_p)]


class Driver:
    """One stream of one library; feed() gives it input pieces and output room and records what every call did."""

    def __init__(self, lib, prefix, fmt, deflate):
        self.lib, self.deflate = lib, deflate
        self.s = Stream()
        self.name = prefix + ("deflate" if deflate else "inflate")
        init = getattr(lib, self.name + "_init")
        init.restype = ctypes.c_int
        self.rc_init = init(ctypes.c_int(fmt), ctypes.byref(self.s)) if prefix == "ms_" else init(ctypes.byref(self.s))
        self.f = getattr(lib, self.name)
        self.f.restype = ctypes.c_int
        self.end = getattr(lib, self.name + "_end")
        self.end.restype = ctypes.c_int
        self.out = bytearray()
        self.pending = b""

    def call(self, piece, cap, flush):
        data = self.pending + piece
        ibuf = ctypes.create_string_buffer(data, max(len(data), 1))
        obuf = ctypes.create_string_buffer(max(cap, 1) + 2)
        self.s.in_ = ctypes.addressof(ibuf); self.s.in_avail = len(data)
        self.s.out = ctypes.addressof(obuf); self.s.out_avail = cap
        rc = self.f(ctypes.byref(self.s), ctypes.c_int(flush)) if self.deflate else self.f(ctypes.byref(self.s))
        used, made = len(data) - self.s.in_avail, cap - self.s.out_avail
        self.out += obuf.raw[:made]
        self.pending = data[used:]
        return rc, used, made

    def finish(self):
        self.s.in_ = None; self.s.in_avail = 0; self.s.out = None; self.s.out_avail = 0
        return self.end(ctypes.byref(self.s))


def both(ms, reflib, prefix, fmt, deflate):
    return Driver(ms.lib(), prefix, fmt, deflate), Driver(reflib.lib, prefix, fmt, deflate)


def drive(ms, reflib, prefix, fmt, deflate, data, rnd, max_piece, max_cap, use_flush):
    a, b = both(ms, reflib, prefix, fmt, deflate)
    assert a.rc_init == b.rc_init == 0
    pos, trace = 0, []
    for step in range(60000):
        last = pos >= len(data)
        piece = data[pos:pos + rnd.randint(0, max_piece)]
        pos += len(piece)
        cap = rnd.choice((0, 1, 2, 3, rnd.randint(0, max_cap), max_cap))
        flush = FINISH if (pos >= len(data)) else (FLUSH if (use_flush and rnd.random() < 0.1) else NO_FLUSH)
        ra, rb = a.call(piece, cap, flush), b.call(piece, cap, flush)
        trace.append((len(piece), cap, flush, ra, rb))
        assert ra == rb, trace[-5:]
        if ra[0] < 0 or ra[0] == STREAM_END:
            break
        if last and not piece and cap > 0 and ra[1] == 0 and ra[2] == 0:
            break                                         # nothing more to give and no progress: the stream is as far as it gets
    else:
        raise AssertionError("stream did not finish")
    assert bytes(a.out) == bytes(b.out)
    assert (a.s.in_total, a.s.out_total) == (b.s.in_total, b.s.out_total)
    assert a.finish() == b.finish()
    return bytes(a.out), ra[0]


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
