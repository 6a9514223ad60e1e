"""TEST INFRASTRUCTURE ONLY -- compile oracle/l2s_oracle.c (gcc, OpenMP) into oracle/_build/libl2s_oracle.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "l2s_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libl2s_oracle.so")


def build(force=False):
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c99", "-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
