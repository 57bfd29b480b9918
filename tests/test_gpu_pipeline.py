"""Pipelined host wrappers (mscomp_shim.cu: lznt1_compress slices, lznt1_decompress ranges with the GPU chunk index) and the
public index call.  The wrappers read their thresholds once per process, so the many-slice runs happen in a subprocess
with small MSCOMP_PIPE_* values (tests/pipe_check.py); the default thresholds are covered by test_gpu_fullsize.py."""
import os
import subprocess
import sys

import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env", [
    {"MSCOMP_PIPE_MIN": "65536", "MSCOMP_PIPE_RANGE": "131072", "MSCOMP_PIPE_SLICE": "262144"},
    {"MSCOMP_PIPE_MIN": "65536", "MSCOMP_PIPE_RANGE": "65536", "MSCOMP_PIPE_SLICE": "65536"},
    {"MSCOMP_PIPE_MIN": "1048576", "MSCOMP_PIPE_RANGE": "1048576", "MSCOMP_PIPE_SLICE": "1048576"},
], ids=["range128k", "range64k", "range1m"])
def test_pipelined_wrappers_match_oracle(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "pipe_check.py")], env=e, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "pipe_check ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def host_walk(c):
    offs, ip, n = [], 0, len(c)
    while n - ip >= 2:
        h = c[ip] | (c[ip + 1] << 8)
        if h == 0:
            return offs, (3 if n - ip == 2 else 4), ip
        sz = (h & 0xFFF) + 3
        if sz > n - ip:
            return offs, 5, ip
        if (h & 0x7000) != 0x3000:
            return offs, 6, ip
        offs.append(ip)
        ip += sz
    return offs, (0 if n - ip == 0 else (1 if c[ip] == 0 else 2)), ip


def test_gpu_chunk_index_equals_the_header_walk(ms, oracle):
    import numpy as np
    import torch
    ctx = ms.GpuContext()
    plain = ms.generate("text", 6 << 20, 5) + ms.generate("random", 300000, 6) + ms.generate("runs", 1 << 20, 7)
    comp = oracle.compress("lznt1", plain)
    for c in (comp, comp + b"\0", comp + b"\x33", comp + b"\0\0", comp + b"\0\0\x01", comp[:-5], comp[:len(comp) // 3],
              comp[:70000] + b"\0\0" + comp[70002:], comp[:200000] + bytes([comp[200000] ^ 0x5A]) + comp[200001:], b"", b"\0", b"\x01"):
        d = torch.from_numpy(np.frombuffer(c + b"\0" * 16, dtype=np.uint8).copy()).cuda()
        d_off, info = ctx.lznt1_index(d, in_len=len(c))
        offs, tail, ip = host_walk(c)
        assert (info["nchunks"], info["tail"], info["consumed"], info["status"]) == (len(offs), tail, ip, 0), len(c)
        assert d_off[:len(offs)].cpu().tolist() == offs
    # too small an offset array: the count is still reported
    d = torch.from_numpy(np.frombuffer(comp, dtype=np.uint8).copy()).cuda()
    _, info = ctx.lznt1_index(d, chunk_cap=10)
    assert info["status"] == -5 and info["nchunks"] == len(host_walk(comp)[0])
