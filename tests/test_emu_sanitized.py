"""All kernels, compiled for the CPU with -fsanitize=address,undefined (shared-memory indexing, alignment, overflow)."""
import glob
import os
import subprocess
import sys

import pytest
from conftest import ROOT


def test_kernels_clean_under_asan_ubsan(oracle, ms):
    libs = []
    for name in ("libasan.so", "libubsan.so"):
        hits = sorted(glob.glob("/usr/lib/gcc/x86_64-linux-gnu/*/" + name))
        if not hits:
            pytest.skip("sanitizer runtime not installed")
        libs.append(hits[-1])
    d = os.path.join(ROOT, "tests", "emu")
    r = subprocess.run(["make", "-s", "-C", d, "libemu_san.so"], capture_output=True, text=True)
    if r.returncode:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-200:])
    env = dict(os.environ, LD_PRELOAD=" ".join(libs), ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_sanitize_run.py")], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "sanitized run ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
