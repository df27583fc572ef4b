"""Builds the in-tree CUDA shared library (sm_100a only) with nvcc.  `python -m dysturbd_b200._build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "dys_api.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "dys_common.cuh"), os.path.join(HERE, "csrc", "dys_day.cuh"),
        os.path.join(HERE, "csrc", "dys_load.cuh"), os.path.join(HERE, "csrc", "dys_routine.cuh"), os.path.join(HERE, "csrc", "dys_network.cuh"),
        os.path.join(HERE, "csrc", "philox.cuh"),
        os.path.join(HERE, "..", "include", "dysturbd.h")]
LIB = os.path.join(HERE, "libdysturbd.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-ccbin", "g++"]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


LIB_CHECKED = os.path.join(HERE, "libdysturbd_checked.so")


def build(force=False, verbose=False, checked=False):
    """checked=True: the same library with DYS_CHECKS=1 (bounds checks on every index the day kernels form), built next
    to the product library and selected with DYS_LIB=<path> (tests/test_gpu_checked.py)"""
    lib = LIB_CHECKED if checked else LIB
    if not force and os.path.exists(lib) and not any(os.path.exists(d) and os.path.getmtime(d) > os.path.getmtime(lib) for d in DEPS):
        return lib
    cmd = [nvcc()] + NVCC_FLAGS + (["-DDYS_CHECKS=1"] if checked else []) + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", lib]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0 or "warning #" in r.stdout:      # (e.g. #940 missing return: never build silently over it)
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % lib)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, checked="--checked" in sys.argv))
