"""Build libpwr_b200.so in-tree with nvcc for sm_100a (no torch/pybind linkage).

    python -m pixelwiseregression_b200.build [--force]

The library is a plain C-ABI shared object (include/pwr.h) loaded with ctypes,
so it travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpwr_b200.so")
SOURCES = ["decoder.cu", "sfr.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "pwr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--cudart", "static", "-Xptxas", "-v"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; libpwr_b200.so cannot be built")


STAMP = LIB + ".stamp"


def source_digest():
    """sha256 over the sources, headers and flags the library is built from.  The stamp written next to
    the .so travels with it (gpurun snapshot), so a stale library is detected by content, not by mtimes
    (which a snapshot copy does not preserve)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.isfile(LIB) or not os.path.isfile(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_digest()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpwr_b200.so")
    if verbose:
        print(res.stderr)
    with open(os.path.join(PKG, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_digest() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
