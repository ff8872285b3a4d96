"""Build recipe for the C oracle (TEST INFRASTRUCTURE ONLY): gcc -> oracle/_build/libgsr_oracle.so.

-ffp-contract=off keeps every float operation a single rounded op, as the restatement requires
(oracle/gsr_oracle.c header).  The reference itself is Python, so there is no ``oracle/_ref`` binary:
it cannot be compiled and cannot travel to the GPU box (DESIGN.md "Oracle").
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gsr_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libgsr_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-Wextra", SRC, "-o", OUT, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
