"""Shared fixtures.  `-m "not gpu"`: oracle vs golden vectors, host logic, emulated kernels, ABI surface.
`-m gpu`: parity of the CUDA path (through the C ABI) against the oracle."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


MAXSZ = {
    "lznt1": lambda n: n + 3 + 2 * ((n + 4095) // 4096),
    "xpress": lambda n: n + 4 + 4 * (n // 32),
    "xpress_huff": lambda n: n + 34 + 258 + 258 * (n // 65536),
}
FORMATS = ("lznt1", "xpress", "xpress_huff")


class CodecLib:
    """Uniform (status, bytes) interface over the oracle restatement or the compiled reference."""

    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix

    def call(self, name, data, cap):
        out = ctypes.create_string_buffer(max(cap, 1))
        n = ctypes.c_size_t(cap)
        f = getattr(self.lib, self.prefix + name)
        f.restype = ctypes.c_int
        rc = f(bytes(data), ctypes.c_size_t(len(data)), out, ctypes.byref(n))
        return rc, (out.raw[:n.value] if rc == 0 else None)

    def compress(self, fmt, data):
        rc, c = self.call(fmt + "_compress", data, MAXSZ[fmt](len(data)) + 16)
        assert rc == 0, (fmt, rc)
        return c

    def decompress(self, fmt, data, cap):
        return self.call(fmt + "_decompress", data, cap)


@pytest.fixture(scope="session")
def oracle():
    """oracle/liboracle.so -- our plain-C restatement of the reference (built on demand with gcc)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return CodecLib(os.path.join(ROOT, "oracle", "liboracle.so"), "orc_")


@pytest.fixture(scope="session")
def reflib():
    """oracle/_ref/libMSCompression.so -- the compiled, unmodified reference (skip where it was not built)."""
    p = os.path.join(ROOT, "oracle", "_ref", "libMSCompression.so")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference/src"):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
        else:
            pytest.skip("compiled reference not available")
    return CodecLib(p, "")


@pytest.fixture(scope="session")
def ms():
    import mscompress_b200
    if not os.path.exists(mscompress_b200.GEN_PATH) or not os.path.exists(mscompress_b200.LIB_PATH):
        mscompress_b200.build()
    return mscompress_b200


@pytest.fixture(scope="session")
def emu():
    """tests/emu/libemu.so -- the CUDA kernel sources compiled for the CPU through the SIMT emulator (ASAN+UBSAN
    when the sanitizer runtime can be preloaded, plain otherwise)."""
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d, "fast"], check=True)
    return ctypes.CDLL(os.path.join(d, "libemu.so"))


def golden_path(name):
    return os.path.join(GOLDEN, name)
