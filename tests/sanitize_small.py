"""Small end-to-end exercise of every kernel, meant to be run under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck  python tests/sanitize_small.py
    compute-sanitizer --tool racecheck python tests/sanitize_small.py
Covers the one-shot C ABI (incl. the pipelined LZNT1 wrappers with small MSCOMP_PIPE_* thresholds), the heavy-chunk path of
the LZNT1 encoder, the GPU chunk index, the indexed plain-Xpress and Xpress-Huffman batch calls, the chunk-parallel one-shot
Xpress-Huffman path (streams made here and by the oracle) and the Xpress stream decoder.  SANITIZE_ONLY=new runs the round-2 additions only."""
import os
import random
import sys

os.environ.setdefault("MSCOMP_PIPE_MIN", "65536")
os.environ.setdefault("MSCOMP_PIPE_RANGE", "131072")
os.environ.setdefault("MSCOMP_PIPE_SLICE", "131072")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mscompress_b200 as ms  # noqa: E402

MAXSZ = {"lznt1": lambda n: n + 3 + 2 * ((n + 4095) // 4096), "xpress": lambda n: n + 4 + 4 * (n // 32),
         "xpress_huff": lambda n: n + 34 + 258 + 258 * (n // 65536)}
rnd = random.Random(1)
extra = [("lznt1", bytes(rnd.choice(b"ACGT") for _ in range(3 * 4096 + 100))),                       # heavy chunks: lazy parse
         ("lznt1", b"".join((b"\x00" * rnd.randint(3, 300)) + bytes([rnd.randint(1, 255)]) for _ in range(200))),
         ("lznt1", ms.generate("text", 400000, 3)),                                                    # several slices / ranges
         ("xpress", ms.generate("text", 300000, 4)), ("xpress", ms.generate("runs", 200000, 5))]       # two-pass encoder, warp parser, block apply
cases = [(fmt, ms.generate(kind, n, 7)) for fmt in ("lznt1", "xpress_huff", "xpress")
         for kind, n in (("text", 70000), ("runs", 20000), ("random", 5000), ("text", 33))] + extra
if os.environ.get("SANITIZE_ONLY") == "new":
    cases = []
for fmt, plain in cases:
    n = len(plain)
    st, comp = ms.raw_call(fmt + "_compress", plain, MAXSZ[fmt](n))
    assert st == 0, (fmt, n, st)
    st, back = ms.raw_call(fmt + "_decompress", comp, n)
    assert st == 0 and back == plain, (fmt, n, st)

ctx = ms.GpuContext()
# GPU chunk index
comp = ms.raw_call("lznt1_compress", ms.generate("text", 300000, 9), MAXSZ["lznt1"](300000))[1]
d = torch.from_numpy(np.frombuffer(comp + b"\0" * 16, dtype=np.uint8).copy()).cuda()
d_off, info = ctx.lznt1_index(d, in_len=len(comp))
assert info["nchunks"] == (300000 + 4095) // 4096 and info["tail"] == 0 and info["consumed"] == len(comp), info
# indexed plain Xpress
plains = [ms.generate("text", 70000, 1), ms.generate("runs", 30000, 2), ms.generate("random", 9000, 3)]
caps = [MAXSZ["xpress"](len(p)) for p in plains]
in_off = np.cumsum([0] + [(len(p) + 63) // 64 * 64 for p in plains])[:-1]
c_off = np.cumsum([0] + [(c + 63) // 64 * 64 for c in caps])[:-1]
blob = np.zeros(int(in_off[-1]) + len(plains[-1]) + 64, dtype=np.uint8)
for p, o in zip(plains, in_off):
    blob[o:o + len(p)] = np.frombuffer(p, dtype=np.uint8)
d_in = torch.from_numpy(blob).cuda()
d_comp = torch.zeros(int(c_off[-1]) + caps[-1] + 64, dtype=torch.uint8, device="cuda")
d_out = torch.zeros_like(d_in)
enc, c_len, st1, k1 = ctx.make_batch(in_off.tolist(), [len(p) for p in plains], c_off.tolist(), caps)
stride = 40
d_index = ctx.xpress_index_tensor(len(plains), stride)
ctx.xpress_compress_indexed(d_in, d_comp, enc, 2048, d_index, stride)
torch.cuda.synchronize()
dec, o_len, st2, k2 = ctx.make_batch(c_off.tolist(), c_len.tolist(), in_off.tolist(), [len(p) for p in plains])
ctx.xpress_decompress_indexed(d_comp, d_out, dec, d_index, stride)
torch.cuda.synchronize()
assert st1.tolist() == [0, 0, 0] and st2.tolist() == [0, 0, 0]
for p, o in zip(plains, in_off.tolist()):
    assert bytes(d_out[o:o + len(p)].cpu().numpy()) == p
# ---- round 2 ----------------------------------------------------------------------------------------------------------
import ctypes  # noqa: E402
import subprocess  # noqa: E402
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
orc = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))


def oracle_compress(fmt, data):
    out = ctypes.create_string_buffer(len(data) + len(data) // 8 + 4096)
    m = ctypes.c_size_t(len(out))
    assert getattr(orc, "orc_%s_compress" % fmt)(data, ctypes.c_size_t(len(data)), out, ctypes.byref(m)) == 0
    return out.raw[:m.value]


# one-shot Xpress-Huffman, multi-chunk: chunk-parallel parse; chunks on their own warps (made here) / in order through the ring (oracle-made)
for plain in (ms.generate("text", 300000, 11), ms.generate("runs", 400000, 12)):
    st, comp = ms.raw_call("xpress_huff_compress", plain, MAXSZ["xpress_huff"](len(plain)))
    assert st == 0 and ms.raw_call("xpress_huff_decompress", comp, len(plain)) == (0, plain)
    assert ms.raw_call("xpress_huff_decompress", oracle_compress("xpress_huff", plain), len(plain)) == (0, plain)
# one-shot plain Xpress, ONE stream: flag groups by speculation, restart index, indexed parse, pointer jumping (kernels/xpress_spec.cuh)
for plain in (ms.generate("text", 300000, 13), ms.generate("runs", 400000, 14) + ms.generate("random", 40000, 15)):
    comp = oracle_compress("xpress", plain)
    assert ms.raw_call("xpress_decompress", comp, len(plain)) == (0, plain)
    assert ms.raw_call("xpress_decompress", comp[:-3], len(plain))[0] != 0
# Xpress-Huffman with the encoder's restart index
plains = [ms.generate("text", 140000, 1), ms.generate("runs", 70000, 2), ms.generate("random", 9000, 3)]
caps = [MAXSZ["xpress_huff"](len(p)) + 16 for p in plains]
in_off = np.cumsum([0] + [(len(p) + 63) // 64 * 64 for p in plains])[:-1]
c_off = np.cumsum([0] + [(c + 63) // 64 * 64 for c in caps])[:-1]
blob = np.zeros(int(in_off[-1]) + len(plains[-1]) + 64, dtype=np.uint8)
for p, o in zip(plains, in_off):
    blob[o:o + len(p)] = np.frombuffer(p, dtype=np.uint8)
d_in = torch.from_numpy(blob).cuda()
d_comp = torch.zeros(int(c_off[-1]) + caps[-1] + 64, dtype=torch.uint8, device="cuda")
d_out = torch.zeros_like(d_in)
enc, c_len, st1, k1 = ctx.make_batch(in_off.tolist(), [len(p) for p in plains], c_off.tolist(), caps)
xh = ctx.xhuff_index_alloc([len(p) for p in plains])
ctx.xhuff_compress_indexed(d_in, d_comp, enc, *xh)
torch.cuda.synchronize()
dec, o_len, st2, k2 = ctx.make_batch(c_off.tolist(), c_len.tolist(), in_off.tolist(), [len(p) for p in plains])
ctx.xhuff_decompress_indexed(d_comp, d_out, dec, *xh)
torch.cuda.synchronize()
assert st1.tolist() == [0, 0, 0] and st2.tolist() == [0, 0, 0]
for p, o in zip(plains, in_off.tolist()):
    assert bytes(d_out[o:o + len(p)].cpu().numpy()) == p
# xpress_inflate: the stream decoder with its state in device memory
sys.path.insert(0, os.path.join(ROOT, "tests"))
from stream_util import Driver  # noqa: E402
plain = ms.generate("text", 50000, 5)
comp = oracle_compress("xpress", plain)
d = Driver(ms.lib(), "xpress_", False, None)
pos = 0
for _ in range(10000):
    if pos >= len(comp) and len(d.out) >= len(plain):
        break
    piece = comp[pos:pos + 3001]; pos += len(piece)
    d.call(piece, 4099, 0)
assert bytes(d.out) == plain and d.finish() == 0
print("sanitize_small ok")
