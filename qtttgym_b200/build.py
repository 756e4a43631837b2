"""Builds qtttgym_b200/csrc/libqttt_b200.so for sm_100a with nvcc (in-tree, so the .so travels
to the GPU box with the repo snapshot).  ``python -m qtttgym_b200.build [--force]``"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = [os.path.join(CSRC, "qttt_kernels.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "qttt_core.cuh"), os.path.join(CSRC, "qttt_mcts.cuh"),
                  os.path.join(os.path.dirname(CSRC), "..", "include", "qttt_b200.h")]
# QTTT_B200_LIB: load another build of the library instead (kernel experiments only)
LIB = os.environ.get("QTTT_B200_LIB") or os.path.join(CSRC, "libqttt_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--cudart", "shared"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libqttt_b200.so")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
