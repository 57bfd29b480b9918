"""Engines for the acceptance harness -- a Python 3 restatement of the reference's test/compressors.py for the engines that
exist on a Linux box: every engine is `OpenSrc` (the reference's ctypes class, test/compressors.py:214-314) over one
libMSCompression:

    B200       this repository's library (CUDA kernels behind the same C ABI)
    Reference  the unmodified reference compiled by oracle/Makefile into oracle/_ref/ (where it was built)

The Windows engines of the reference harness (RtlCompressBuffer, the .NET / cabinet wrappers; test/compressors.py:60-186) have no
counterpart here.  Dictionaries map engine name -> codec object, one dictionary per format, as upstream (:311-314).
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mscompress_b200 as ms                                     # noqa: E402
from mscompress_b200 import CompressionFormat, OpenSrc            # noqa: E402

FORMATS = {"none": CompressionFormat.NoCompression, "lznt1": CompressionFormat.LZNT1, "xpress": CompressionFormat.Xpress,
           "xpress-huffman": CompressionFormat.XpressHuffman}


def load_libraries(need_gpu=True):
    """name -> loaded library.  The B200 engine needs a CUDA device (it has no CPU path); the reference engine needs oracle/_ref."""
    libs = {}
    if need_gpu:
        libs["B200"] = ms.lib()
    ref = os.path.join(ROOT, "oracle", "_ref", "libMSCompression.so")
    if os.path.exists(ref):
        libs["Reference"] = ctypes.CDLL(ref)
    return libs


def engines_for(fmt, need_gpu=True):
    """The reference harness' per-format dictionary (test/compressors.py:311-314): engine name -> codec object."""
    if fmt.lower() not in FORMATS:
        return None
    return {name: OpenSrc(FORMATS[fmt.lower()], library) for name, library in load_libraries(need_gpu).items()}
