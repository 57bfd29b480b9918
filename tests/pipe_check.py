"""Run in a subprocess by tests/test_gpu_pipeline.py with small MSCOMP_PIPE_* thresholds, so that the pipelined host
wrappers of lznt1_compress / lznt1_decompress (mscomp_shim.cu) run many slices / ranges on a few MiB.  Every result is
compared with the oracle (status and bytes)."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import subprocess  # noqa: E402

import mscompress_b200 as ms  # noqa: E402
from conftest import MAXSZ, ROOT, CodecLib  # noqa: E402

subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
oracle = CodecLib(os.path.join(ROOT, "oracle", "liboracle.so"), "orc_")
rnd = random.Random(21)
checked = 0


import ctypes  # noqa: E402

fallbacks = ms.lib().msc_pipeline_fallbacks
fallbacks.restype = ctypes.c_ulonglong


def same(src, cap):
    global checked
    got, want = ms.raw_call("lznt1_decompress", src, cap), oracle.decompress("lznt1", src, cap)
    assert got == want, ("decompress", len(src), cap, got[0], want[0], len(got[1]), len(want[1]))
    checked += 1


for kind, n in (("text", 3 << 20), ("random", (1 << 20) + 77), ("runs", 2 << 20), ("text", (5 << 20) + 4097)):
    plain = ms.generate(kind, n, 3 + n % 7)
    comp = oracle.compress("lznt1", plain)
    assert ms.raw_call("lznt1_compress", plain, MAXSZ["lznt1"](n)) == (0, comp), ("compress", kind, n)
    assert ms.raw_call("lznt1_compress", plain, len(comp)) == (0, comp)
    assert ms.raw_call("lznt1_compress", plain, len(comp) - 1)[0] == oracle.call("lznt1_compress", plain, len(comp) - 1)[0]
    f0 = fallbacks()
    for cap in (n, n + 1, n + 5000, 4 * n):
        same(comp, cap)
    assert fallbacks() == f0, "a well-formed stream must stay on the pipelined path"
    for cap in (n - 1, n // 2):
        same(comp, cap)
    for tail in (b"\0", b"\x55", b"\0\0", b"\0\0\x01", b"\0" * 9):
        same(comp + tail, n)
        same(comp + tail, n + 10)
    for cut in (1, 2, 3, 1000, len(comp) // 2):
        same(comp[:-cut], n)
    for _ in range(6):                                     # damaged payloads and headers
        b = bytearray(comp)
        for _ in range(rnd.randint(1, 3)):
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        same(bytes(b), n + rnd.choice((0, 64)))

# short chunks in mid-stream (streams cut by a flushing compressor), at several places of a long stream
parts = [ms.generate("text", sz, 40 + i) for i, sz in enumerate((700000, 100, 900000, 4095, 1, 800000, 50))]
stream = b"".join(oracle.compress("lznt1", x) for x in parts)
total = sum(len(x) for x in parts)
for cap in (total, total + 100, total - 1):
    same(stream, cap)
print("pipe_check ok", checked)
