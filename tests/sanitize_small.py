"""Small end-to-end exercise of every kernel, meant to be run under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck  python tests/sanitize_small.py
    compute-sanitizer --tool racecheck python tests/sanitize_small.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mscompress_b200 as ms  # noqa: E402

MAXSZ = {"lznt1": lambda n: n + 3 + 2 * ((n + 4095) // 4096), "xpress": lambda n: n + 4 + 4 * (n // 32),
         "xpress_huff": lambda n: n + 34 + 258 + 258 * (n // 65536)}
for fmt in ("lznt1", "xpress_huff", "xpress"):
    for kind, n in (("text", 70000), ("runs", 20000), ("random", 5000), ("text", 33)):
        plain = ms.generate(kind, n, 7)
        st, comp = ms.raw_call(fmt + "_compress", plain, MAXSZ[fmt](n))
        assert st == 0, (fmt, kind, n, st)
        st, back = ms.raw_call(fmt + "_decompress", comp, n)
        assert st == 0 and back == plain, (fmt, kind, n, st)
print("sanitize_small ok")
