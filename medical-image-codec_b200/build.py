"""Builds libmicgpu.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmicgpu.so")
SOURCES = ["k_table.cu", "k_ans.cu", "k_ans_serial.cu", "k_huff.cu", "k_rle.cu", "k_delta.cu", "k_delta_scan.cu", "k_grad.cu", "k_misc.cu", "k_wavelet.cu", "k_enc_rle.cu", "k_enc_fse.cu", "k_enc_front.cu", "micgpu_host.cu", "micgpu_enc_host.cu", "micgpu_huff_host.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared", "-t", "0",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "micgpu.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("MICGPU_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
