"""Builds the oracle's C restatement (oracle/so_oracle.c) into oracle/_build/libso_oracle.so with gcc.
The reference itself is Python 2 + ASE + zacros_wrapper and cannot be compiled or imported here, so there is no
oracle/_ref; this library is the oracle *port*, test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "so_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libso_oracle.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
