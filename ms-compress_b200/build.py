"""Build the native libraries of the package IN-TREE (ms-compress_b200/lib/).

  libMSCompression.so -- CUDA kernels (sm_100a) + C-ABI shim, built with nvcc; cudart linked statically, no torch.
  libmsdatagen.so     -- host-side synthetic data generators used by bench.py and the tests (plain C).

nvcc cross-compiles without a GPU, so this runs in the build container as well as on the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]
CU_SOURCES = ["mscomp_gpu.cu", "mscomp_shim.cu", "mscomp_stream.cu", "mscomp_multi.cu"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _all_sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    inc = os.path.join(os.path.dirname(HERE), "include")
    for root, _, files in os.walk(inc):
        out += [os.path.join(root, f) for f in files]
    return out


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    lib = os.path.join(LIB, "libMSCompression.so")
    if force or _newer(lib, _all_sources()):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + [os.path.join(CSRC, s) for s in CU_SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building libMSCompression.so")
    gen = os.path.join(LIB, "libmsdatagen.so")
    gsrc = os.path.join(CSRC, "datagen.c")
    if force or _newer(gen, [gsrc]):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", gen, gsrc], check=True)
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
