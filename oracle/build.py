"""TEST INFRASTRUCTURE ONLY -- builds the C oracle: oracle/qttt_oracle.c -> oracle/_build/libqttt_oracle.so.

There is no ``oracle/_ref``: the reference is pure Python (no C/C++ sources to compile), so
the "real reference" leg is the live import in ``oracle/refload.py`` (build container only)
and the recorded fixtures in ``tests/golden/``.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "qttt_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libqttt_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-Wall", "-Wextra",
           "-o", OUT, SRC]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
