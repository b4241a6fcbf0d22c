"""Builds the in-tree native libraries.

``libchaorec_b200.so``  hand-written sm_100a CUDA kernels + the C ABI of include/chaorec_b200.h
                        (nvcc cross-compiles without a GPU; the .so is git-ignored but travels
                        to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libchaorec_b200.so"
STAMP = PKG / ".libchaorec_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in _sources() + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "chaorec_b200.h"]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + [str(s) for s in _sources()]
    res = subprocess.run(cmd, cwd=str(ROOT), capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libchaorec_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
