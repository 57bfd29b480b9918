"""The C-ABI library loads and exports every symbol the public headers declare (no compute calls -- no GPU here)."""
import ctypes
import glob
import os
import re

from conftest import ROOT


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"MSCOMPAPI\s+[\w\s\*]+?\b(\w+)\s*\(", text):
            names.add(m.group(1))
    return names


# The reference's exported C symbols (`nm -D` of the reference build): the drop-in boundary, 30 names.
REFERENCE_SYMBOLS = (
    ["ms_compress", "ms_decompress", "ms_max_compressed_size"]
    + ["ms_%s%s" % (d, s) for d in ("deflate", "inflate") for s in ("_init", "", "_end")]
    + ["lznt1_compress", "lznt1_decompress", "lznt1_max_compressed_size"]
    + ["lznt1_%s%s" % (d, s) for d in ("deflate", "inflate") for s in ("_init", "", "_end")]
    + ["xpress_compress", "xpress_decompress", "xpress_max_compressed_size"]
    + ["xpress_%s%s" % (d, s) for d in ("deflate", "inflate") for s in ("_init", "", "_end")]
    + ["xpress_huff_compress", "xpress_huff_decompress", "xpress_huff_max_compressed_size"]
)


def _exported(path):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if " T " in ln}


def test_library_exports_the_reference_boundary(ms):
    """`nm -D` of our library is a superset of the reference's 30 C symbols -- the list is fixed here, not derived from our headers."""
    assert len(REFERENCE_SYMBOLS) == 30 and len(set(REFERENCE_SYMBOLS)) == 30
    ours = _exported(ms.LIB_PATH)
    missing = sorted(set(REFERENCE_SYMBOLS) - ours)
    assert not missing, missing
    assert set(REFERENCE_SYMBOLS) <= declared_symbols()             # and every one of them is declared in include/*.h
    ref = os.path.join(ROOT, "oracle", "_ref", "libMSCompression.so")
    if os.path.exists(ref):                                          # the compiled reference, where it was built
        ref_c = {n for n in _exported(ref) if not n.startswith("_Z")}
        assert ref_c == set(REFERENCE_SYMBOLS), sorted(ref_c ^ set(REFERENCE_SYMBOLS))
        assert not (ref_c - ours)


def test_library_exports_declared_symbols(ms):
    lib = ctypes.CDLL(ms.LIB_PATH)
    names = declared_symbols()
    assert {"ms_compress", "ms_decompress", "ms_max_compressed_size", "lznt1_compress", "lznt1_decompress"} <= names
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_max_compressed_size_matches_reference_formulas(ms):
    L = ms.lib()
    # reference values: SURVEY.md 8(b) table, src/lznt1_compress.cpp:27, src/xpress_compress.cpp:30, src/xpress_huff_compress.cpp:46
    assert L.ms_max_compressed_size(2, 65536) == 65571
    assert L.ms_max_compressed_size(3, 65536) == 73732
    assert L.ms_max_compressed_size(4, 65536) == 66086
    assert L.ms_max_compressed_size(0, 65536) == 65536
    assert L.ms_max_compressed_size(1, 65536) == 2 ** 64 - 1
    assert L.ms_max_compressed_size(5, 65536) == 2 ** 64 - 1


def test_bad_format_is_arg_error(ms):
    out = ctypes.create_string_buffer(16)
    n = ctypes.c_size_t(16)
    for fmt in (1, 5, 99):
        assert ms.lib().ms_compress(fmt, b"abc", 3, out, ctypes.byref(n)) == ms.MSCOMP_ARG_ERROR
        assert ms.lib().ms_decompress(fmt, b"abc", 3, out, ctypes.byref(n)) == ms.MSCOMP_ARG_ERROR


def test_none_format_copies(ms):
    out = ctypes.create_string_buffer(16)
    n = ctypes.c_size_t(16)
    assert ms.lib().ms_compress(0, b"hello", 5, out, ctypes.byref(n)) == 0 and out.raw[:n.value] == b"hello"
    n = ctypes.c_size_t(2)
    assert ms.lib().ms_decompress(0, b"hello", 5, out, ctypes.byref(n)) == ms.MSCOMP_BUF_ERROR


def test_public_headers_compile_as_c_and_cpp(tmp_path):
    """include/*.h is what a C (or C++) program of the reference's users sees: plain C99, no CUDA or torch types."""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    for h in ("mscomp.h", "lznt1.h", "xpress.h", "xpress_huff.h", "mscomp_gpu.h"):
        for cc, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "c++")):
            if not shutil.which(cc):
                continue
            src = '#include "%s"\nint main(void) { mscomp_stream s; (void)s; return (int)sizeof(MSCompStatus) - (int)sizeof(MSCompStatus); }\n' % h
            r = subprocess.run([cc, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, "-x", lang, "-"],
                               input=src, capture_output=True, text=True)
            assert r.returncode == 0, (h, cc, r.stderr)


def test_codec_calls_fail_loudly_without_a_gpu(ms):
    """No CPU implementation hides behind the C ABI: without a CUDA device every codec entry point returns MSCOMP_MEM_ERROR."""
    import torch
    if torch.cuda.is_available():
        return                                            # (the GPU box runs the real thing)
    data = bytes(5000)
    for fmt in ("lznt1", "xpress", "xpress_huff"):
        assert ms.raw_call(fmt + "_compress", data, 10000)[0] == -4
        assert ms.raw_call(fmt + "_decompress", data, 10000)[0] in (-4, -3, -5)      # (argument / header checks may come first)
    assert ms.raw_call("lznt1_decompress", b"\x00\xb0" + bytes(100), 4096)[0] == -4


def test_shard_range_is_the_python_partition(ms):
    """mscomp_gpu_shard_range (C ABI, no GPU needed) = mscompress_b200.sharding.partition: contiguous, covering, balanced."""
    from importlib import import_module
    sh = import_module("mscompress_b200.sharding")
    L = ms.lib()
    for units in (0, 1, 7, 8, 262144, 65536 * 3 + 5, 2 ** 40 + 123):
        for world in (1, 2, 3, 4, 8):
            total = 0
            for rank in range(world):
                first, count = ctypes.c_uint64(), ctypes.c_uint64()
                L.mscomp_gpu_shard_range(ctypes.c_uint64(units), rank, world, ctypes.byref(first), ctypes.byref(count))
                lo, hi = sh.partition(units, world, rank)
                assert (first.value, first.value + count.value) == (lo, hi)
                total += count.value
            assert total == units


def test_multi_gpu_calls_fail_cleanly_without_a_context(ms):
    L = ms.lib()
    buf = ctypes.create_string_buffer(128)
    assert L.mscomp_gpu_comm_init(None, ctypes.addressof(buf), 0, 1) == ms.MSCOMP_ARG_ERROR
    assert L.mscomp_gpu_sizes_allgather(None, None, 1, None) == ms.MSCOMP_ARG_ERROR
    assert L.mscomp_gpu_comm_size(None) == 1 and L.mscomp_gpu_collectives(None) == 0
