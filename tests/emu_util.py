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


def emu_lznt1_decode(emu, comp, offs=None, grid=3):
    offs = lznt1_walk(comp) if offs is None else offs
    nch = len(offs)
    a, addr = padded(comp)
    off = np.array(offs, dtype=np.uint64)
    out = np.zeros(nch * 4096 + PAD, dtype=np.uint8)
    sz = np.zeros(max(nch, 1), dtype=np.uint32)
    st = np.zeros(max(nch, 1), dtype=np.int32)
    emu.emu_lznt1_decode(ctypes.c_void_p(addr), ctypes.c_uint64(len(comp)), P(off), ctypes.c_uint32(nch), P(out), P(sz), P(st),
                         ctypes.c_uint32(grid))
    chunks = [out[i * 4096:i * 4096 + sz[i]].tobytes() for i in range(nch)]
    return st[:nch], sz[:nch], chunks


def emu_lznt1_encode(emu, data, grid=3):
    n = len(data)
    nch = (n + 4095) // 4096
    a, addr = padded(data)
    slots = np.zeros(nch * 4112 + PAD, dtype=np.uint8)
    ss = np.zeros(nch + 1, dtype=np.uint32)
    offs = np.zeros(nch + 1, dtype=np.uint64)
    out = np.zeros(n + 2 * nch + 2 * PAD, dtype=np.uint8)
    emu.emu_lznt1_encode(ctypes.c_void_p(addr), ctypes.c_uint64(n), ctypes.c_uint32(nch), P(slots), P(ss), P(offs),
                         ctypes.c_void_p(out.ctypes.data + 16), ctypes.c_uint32(grid))
    return out[16:16 + int(offs[nch])].tobytes(), offs
