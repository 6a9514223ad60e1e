"""Build libl2s_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m l2s_b200.build            # rebuild if sources are newer than the library
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libl2s_b200.so")
SOURCES = ["l2s_capi.cu"]
HEADERS = ["l2s_device.cuh", "l2s_kernels.cuh", os.path.join("..", "..", "include", "l2s_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


PYLISTS_SRC = os.path.join(CSRC, "l2s_pylists.c")
PYLISTS_PATH = os.path.join(LIB_DIR, "l2s_pylists.so")


def build_pylists(force=False):
    """CPython glue (list <-> pinned buffer conversions of the reference's list-shaped API); gcc, no CUDA."""
    import sysconfig
    if not force and os.path.isfile(PYLISTS_PATH) and os.path.getmtime(PYLISTS_PATH) >= os.path.getmtime(PYLISTS_SRC):
        return PYLISTS_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], "-o", PYLISTS_PATH, PYLISTS_SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n%s" % res.stderr)
    return PYLISTS_PATH


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources.  Returns the library path."""
    try:
        build_pylists(force)
    except Exception as e:   # optional accelerator of the Python list conversions only
        if verbose:
            print("l2s_pylists not built:", e)
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
