"""Regenerates tests/golden/*.  Run in the BUILD CONTAINER only (needs oracle/_ref/libMSCompression.so, i.e. the
compiled, unmodified reference):   python tests/golden/make_golden.py

Produces
  kat.json        the five MS-XCA documentation vectors, each verified here to decode with
                  the compiled reference, plus the reference encoder's own output for the same plaintexts
  ref_<fmt>_<kind>_<n>.bin   reference-compressed generator data (seed 42) -- small fixtures the GPU box decodes
  digests.json    md5 / sizes of reference-compressed generator data at 64 KiB and 1 MiB (seed 42)
"""
import ctypes
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import mscompress_b200 as ms  # noqa: E402  (only the host-side data generators are used here)

ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libMSCompression.so"))


def call(name, data, cap):
    out = ctypes.create_string_buffer(max(cap, 1))
    n = ctypes.c_size_t(cap)
    f = getattr(ref, name)
    f.restype = ctypes.c_int
    rc = f(data, ctypes.c_size_t(len(data)), out, ctypes.byref(n))
    return rc, (out.raw[:n.value] if rc == 0 else None)


MAXSZ = {"lznt1": lambda n: n + 3 + 2 * ((n + 4095) // 4096), "xpress": lambda n: n + 4 + 4 * (n // 32),
         "xpress_huff": lambda n: n + 34 + 258 + 258 * (n // 65536)}

LZNT1_PLAIN = (b"F# F# G A A G F# E D D E F# F# E E F# F# G A A G F# E D D E F# E D D E E F# D E F# G F# D E F# G F# E D E A "
               b"F# F# G A A G F# E D D E F# E D D\0")
LZNT1_COMP = bytes.fromhex("38B08846232000204720410010A24701A045204400084501507900C045200524138805B4024A44EF0358028C091601484500BE009E000401189000")
ALPHA = bytes(range(97, 123))
ABC100 = b"abc" * 100


def xh_table(entries):
    t = bytearray(256)
    for k, v in entries.items():
        t[k] = v
    return bytes(t)


XH_ALPHA = xh_table({0x30: 0x50, **{i: 0x55 for i in range(0x31, 0x3B)}, 0x3B: 0x45, 0x3C: 0x44, 0x3D: 0x04, 0x80: 0x04}) + \
    bytes.fromhex("d8523ed794115be9195ff9d67cdf8d0400000000")
XH_ABC = xh_table({0x30: 0x30, 0x31: 0x23, 0x80: 0x02, 0x8F: 0x20}) + bytes.fromhex("a8dc0000ff2601")

kats = [
    ("lznt1", "msxca_lznt1_example", LZNT1_PLAIN, LZNT1_COMP),
    ("xpress", "msxca_xpress_alphabet", ALPHA, bytes.fromhex("3f000000") + ALPHA),
    ("xpress", "msxca_xpress_abc100", ABC100, bytes.fromhex("ffffff1f616263170 00fff2601".replace(" ", ""))),
    ("xpress_huff", "msxca_xhuff_alphabet", ALPHA, XH_ALPHA),
    ("xpress_huff", "msxca_xhuff_abc100", ABC100, XH_ABC),
]
out = []
for fmt, name, plain, comp in kats:
    rc, dec = call(fmt + "_decompress", comp, len(plain) + 64)
    assert rc == 0 and dec == plain, (name, rc)
    rc, enc = call(fmt + "_compress", plain, MAXSZ[fmt](len(plain)) + 16)
    assert rc == 0
    out.append({"format": fmt, "name": name, "plain_hex": plain.hex(), "compressed_hex": comp.hex(), "reference_encoder_hex": enc.hex()})
json.dump(out, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

digests = {}
for kind in ("text", "random", "runs"):
    for n in (65536, 1 << 20):
        d = ms.generate(kind, n, 42)
        e = {"input_md5": hashlib.md5(d).hexdigest()}
        for fmt in MAXSZ:
            rc, c = call(fmt + "_compress", d, MAXSZ[fmt](n) + 16)
            assert rc == 0
            e[fmt] = {"size": len(c), "md5": hashlib.md5(c).hexdigest()}
        digests["%s_%d" % (kind, n)] = e
json.dump(digests, open(os.path.join(HERE, "digests.json"), "w"), indent=1)

# small compressed fixtures (decoded on the GPU box where the reference cannot be rebuilt)
for kind, n in (("text", 20000), ("runs", 70000), ("random", 5000), ("text", 70000)):
    d = ms.generate(kind, n, 42)
    for fmt in MAXSZ:
        rc, c = call(fmt + "_compress", d, MAXSZ[fmt](n) + 16)
        assert rc == 0
        open(os.path.join(HERE, "ref_%s_%s_%d.bin" % (fmt, kind, n)), "wb").write(c)
print("golden vectors written to", HERE)
