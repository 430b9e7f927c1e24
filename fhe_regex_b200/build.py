"""Build libfhe_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fhe_regex_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfhe_b200.so")
SOURCES = ["kernels.cu", "br_fused.cu", "br_wide.cu", "br_wide2.cu", "br_duo.cu", "ks_kernels.cu", "ks_umma.cu", "api.cu", "comm.cu", "handles.cu", "keygen.cu", "regex_api.cu", "regex_host.cpp", "client.cpp", "wire.cpp"]
HEADERS = ["br_core.cuh", "br_tmem.cuh", "br_wide.cuh", "br_duo.cuh", "ptx_sync.cuh", "fft32_gen.h", "kernels.h", "context.h", "regex_host.h", os.path.join("..", "..", "include", "fhe_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--shared", "-lpthread", "-ldl",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
