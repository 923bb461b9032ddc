"""Build the C-ABI shared library `libegm_b200.so` in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting `.so`
is git-ignored but travels to the GPU box with the repo snapshot.
Note `-gencode arch=compute_100a,code=sm_100a` (NOT `-arch=sm_100a`, which this toolchain lowers
to compute_100 and then rejects tcgen05).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libegm_b200.so")
SOURCES = ["egm_api.cu", "egm_lowrank.cu", "egm_gemm_tc.cu", "egm_gemm_simt.cu", "egm_kernels.cu", "egm_gpf_fused.cu",
           "egm_error.cu"]
HEADERS = ["egm_gemm.h", "egm_chain.h", "egm_kernels.cuh", "egm_ptx.cuh", os.path.join("..", "..", "include", "egm_b200.h")]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
         "-shared", "-Xcompiler", "-fPIC"] + os.environ.get("EGM_NVCC_FLAGS", "").split()   # e.g. -DEGM_GPF_PROFILE


def _fingerprint() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed since the last build; return the library path."""
    stamp = LIB + ".stamp"
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == fp:
                return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
