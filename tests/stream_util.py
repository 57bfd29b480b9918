"""Call-by-call differential driver for the streaming API: two libraries get the same random schedule of input pieces, output
capacities and flush flags and must agree on status, bytes consumed and bytes produced in EVERY call (tests/test_gpu_streaming.py:
product on the GPU vs the compiled reference; tests/test_stream_logic.py: the product's state machine over the oracle's chunk codec
vs the compiled reference, on the CPU)."""
import ctypes

NO_FLUSH, FLUSH, FINISH = 0, 2, 4
STREAM_END, POSSIBLE_END = 1, 2


class Stream(ctypes.Structure):
    _fields_ = [("format", ctypes.c_int), ("compressing", ctypes.c_bool), ("in_", ctypes.c_void_p), ("in_avail", ctypes.c_size_t),
                ("in_total", ctypes.c_size_t), ("out", ctypes.c_void_p), ("out_avail", ctypes.c_size_t), ("out_total", ctypes.c_size_t),
                ("error", ctypes.c_char * 256), ("warning", ctypes.c_char * 256), ("state", ctypes.c_void_p)]


class Driver:
    """One stream of one library; call() gives it an input piece and output room and reports what the call did."""

    def __init__(self, lib, stem, deflate, fmt=None):
        """stem: "lznt1_" / "ms_" / "hs_lznt1_" ...; fmt: format number for the ms_*_init form, else None."""
        self.lib, self.deflate = lib, deflate
        self.s = Stream()
        name = stem + ("deflate" if deflate else "inflate")
        init = getattr(lib, name + "_init")
        init.restype = ctypes.c_int
        self.rc_init = init(ctypes.c_int(fmt), ctypes.byref(self.s)) if fmt is not None else init(ctypes.byref(self.s))
        self.f = getattr(lib, name)
        self.f.restype = ctypes.c_int
        self.end = getattr(lib, name + "_end")
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


def drive_pair(a, b, deflate, data, rnd, max_piece, max_cap, use_flush):
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


