"""Helpers to drive the emulated kernels (tests/emu/libemu.so) with numpy buffers."""
import ctypes

import numpy as np

PAD = 64


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def padded(data):
    """numpy copy of `data` with PAD readable bytes on both sides; returns (array, address of byte 0)."""
    n = len(data)
    a = np.zeros(n + 2 * PAD, dtype=np.uint8)
    a[PAD:PAD + n] = np.frombuffer(bytes(data), dtype=np.uint8)
    return a, a.ctypes.data + PAD


def lznt1_walk(c):
    offs, p = [], 0
    while p + 2 <= len(c):
        h = c[p] | (c[p + 1] << 8)
        if h == 0:
            break
        offs.append(p)
        p += (h & 0xFFF) + 3
    return offs


def emu_lznt1_decode(emu, comp, offs=None, grid=3, kernel="emu_lznt1_decode"):
    offs = lznt1_walk(comp) if offs is None else offs
    nch = len(offs)
    a, addr = padded(comp)
    off = np.array(offs, dtype=np.uint64)
    out = np.zeros(nch * 4096 + PAD, dtype=np.uint8)
    sz = np.zeros(max(nch, 1), dtype=np.uint32)
    st = np.zeros(max(nch, 1), dtype=np.int32)
    getattr(emu, kernel)(ctypes.c_void_p(addr), ctypes.c_uint64(len(comp)), P(off), ctypes.c_uint32(nch), P(out), P(sz), P(st),
                         ctypes.c_uint32(grid))
    chunks = [out[i * 4096:i * 4096 + sz[i]].tobytes() for i in range(nch)]
    return st[:nch], sz[:nch], chunks


def emu_lznt1_encode(emu, data, grid=3, heavy_work=32768):
    n = len(data)
    nch = (n + 4095) // 4096
    a, addr = padded(data)
    slots = np.zeros(nch * 4112 + PAD, dtype=np.uint8)
    ss = np.zeros(nch + 1, dtype=np.uint32)
    offs = np.zeros(nch + 1, dtype=np.uint64)
    out = np.zeros(n + 2 * nch + 2 * PAD, dtype=np.uint8)
    emu.emu_lznt1_encode(ctypes.c_void_p(addr), ctypes.c_uint64(n), ctypes.c_uint32(nch), P(slots), P(ss), P(offs),
                         ctypes.c_void_p(out.ctypes.data + 16), ctypes.c_uint32(grid), ctypes.c_uint32(heavy_work))
    return out[16:16 + int(offs[nch])].tobytes(), offs


def emu_batch(emu, fn_name, bufs, caps, grid=2, extra=()):
    """Run a batch kernel (decode or encode) over independent buffers; returns (status, out_len, [outputs]).
    extra: further ctypes arguments of the emulator entry point (e.g. the Xpress restart index)."""
    count = len(bufs)
    in_off, pos = [], 0
    for b in bufs:
        in_off.append(pos)
        pos += (len(b) + 7) // 8 * 8 + 8
    blob = bytearray(pos + 8)
    for b, o in zip(bufs, in_off):
        blob[o:o + len(b)] = b
    a, addr = padded(bytes(blob))
    out_off, opos = [], 0
    for c in caps:
        out_off.append(opos)
        opos += (c + 15) // 16 * 16 + 16
    out = np.zeros(opos + 2 * PAD, dtype=np.uint8)
    io = np.array(in_off, dtype=np.uint64)
    il = np.array([len(b) for b in bufs], dtype=np.uint32)
    oo = np.array(out_off, dtype=np.uint64)
    oc = np.array(caps, dtype=np.uint32)
    ol = np.zeros(count, dtype=np.uint32)
    st = np.zeros(count, dtype=np.int32)
    fn = getattr(emu, fn_name)
    fn.restype = ctypes.c_uint32
    emu_batch.last_return = fn(ctypes.c_void_p(addr), ctypes.c_void_p(out.ctypes.data + PAD), ctypes.c_uint32(count), P(io), P(il), P(oo), P(oc),
                               P(ol), P(st), ctypes.c_uint32(grid), *extra)
    outs = [out[PAD + out_off[i]:PAD + out_off[i] + ol[i]].tobytes() for i in range(count)]
    return st, ol, outs


def emu_xhuff_stream_exact(emu, data, grid=2):
    """The one-shot layout of xpress_huff_compress: ONE stream, its 64 KiB chunks are the units of the batch (contiguous input, eos_mode 1);
    returns (status per chunk, packed stream)."""
    n = len(data)
    count = (n + 65535) // 65536
    a, addr = padded(data)
    stride = 65536 + 256 + 64
    out = np.zeros(count * stride + 2 * PAD, dtype=np.uint8)
    io = np.arange(count, dtype=np.uint64) * np.uint64(65536)
    il = np.array([min(65536, n - i * 65536) for i in range(count)], dtype=np.uint32)
    oo = np.arange(count, dtype=np.uint64) * np.uint64(stride)
    oc = np.full(count, stride, dtype=np.uint32)
    ol = np.zeros(count, dtype=np.uint32)
    st = np.zeros(count, dtype=np.int32)
    emu.emu_xhuff_encode_exact(ctypes.c_void_p(addr), ctypes.c_void_p(out.ctypes.data + PAD), ctypes.c_uint32(count), P(io), P(il), P(oo), P(oc), P(ol), P(st),
                               ctypes.c_uint32(grid), ctypes.c_uint32(1))
    packed = b"".join(out[PAD + i * stride:PAD + i * stride + int(ol[i])].tobytes() for i in range(count))
    return st, packed
