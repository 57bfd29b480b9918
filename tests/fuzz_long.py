"""Long randomized differential run against the oracle (not collected by pytest; run by hand on the GPU box):
    MSCOMP_PIPE_MIN=65536 MSCOMP_PIPE_RANGE=131072 MSCOMP_PIPE_SLICE=131072 python tests/fuzz_long.py [seed] [rounds]
Every C-ABI result (status and bytes) must equal the oracle's: LZNT1 byte-identical encode, decode of valid and damaged streams
of all three formats through the one-shot calls (pipelined LZNT1 wrappers, GPU chunk index, warp-per-stream Xpress parser,
block-per-stream apply, lane parsers for Xpress-Huffman)."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import subprocess  # noqa: E402

import mscompress_b200 as ms  # noqa: E402
from conftest import MAXSZ, ROOT, CodecLib  # noqa: E402

subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
oracle = CodecLib(os.path.join(ROOT, "oracle", "liboracle.so"), "orc_")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rnd = random.Random(seed)


def make_plain():
    kind = rnd.choice(("text", "text", "runs", "random", "dna", "zeros", "mix"))
    n = rnd.choice((0, 1, 2, 3, rnd.randint(4, 300), rnd.randint(300, 9000), rnd.randint(9000, 140000), rnd.randint(140000, 700000)))
    if kind in ("text", "runs", "random"):
        return ms.generate(kind, n, rnd.randrange(1 << 30))
    if kind == "dna":
        return bytes(rnd.choice(b"ACGT") for _ in range(min(n, 60000)))
    if kind == "zeros":
        return b"".join((b"\x00" * rnd.randint(1, 400)) + bytes([rnd.randint(0, 255)]) for _ in range(max(1, min(n, 200000) // 200)))
    a = ms.generate("text", n // 2, rnd.randrange(1 << 30))
    return a + ms.generate("runs", n - n // 2, rnd.randrange(1 << 30)) + a[: n // 3]


def damage(c):
    b = bytearray(c)
    if not b:
        return bytes(b)
    mode = rnd.randrange(5)
    if mode == 0:
        for _ in range(rnd.randint(1, 4)):
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
    elif mode == 1:
        del b[rnd.randrange(len(b)):]
    elif mode == 2:
        b += bytes(rnd.getrandbits(8) for _ in range(rnd.randint(1, 5)))
    elif mode == 3:
        p = rnd.randrange(len(b))
        b[p:p + rnd.randint(1, 8)] = bytes(rnd.getrandbits(8) for _ in range(rnd.randint(1, 8)))
    else:
        b += b"\0" * rnd.randint(1, 3)
    return bytes(b)


t0, checks = time.time(), 0
for it in range(rounds):
    plain = make_plain()
    n = len(plain)
    for fmt in ("lznt1", "xpress", "xpress_huff"):
        ref = oracle.compress(fmt, plain)
        st, comp = ms.raw_call(fmt + "_compress", plain, MAXSZ[fmt](n))
        assert st == 0, (it, fmt, n, st)
        if fmt == "lznt1":
            assert comp == ref, (it, "lznt1 encoder differs", n)
        if not (fmt == "xpress" and n == 0):
            assert oracle.decompress(fmt, comp, n) == (0, plain), (it, fmt, "own stream does not decode with the oracle", n)
        for src, cap in ((ref, n), (ref, n + rnd.randint(1, 5000)), (ref, max(0, n - rnd.randint(1, 50))), (damage(ref), n + rnd.choice((0, 0, 64))),
                         (damage(ref), n), (damage(comp), n)):
            got, want = ms.raw_call(fmt + "_decompress", src, cap), oracle.decompress(fmt, src, cap)
            assert got == want, (it, fmt, n, len(src), cap, got[0], want[0])
            checks += 1
    if it % 20 == 0:
        print("round", it, "checks", checks, "%.0fs" % (time.time() - t0), flush=True)
print("fuzz_long ok: seed", seed, "rounds", rounds, "checks", checks)
