"""Build the CPU oracle (test infrastructure) into oracle/liboracle.so.

Links to scipy's bundled OpenBLAS (symbols scipy_d*_) when present so that the CPU baseline uses an
optimised multithreaded BLAS-3, as the reference's SuiteSparse build does; falls back to the naive
loops in chol_oracle.c otherwise.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
SRCS = [os.path.join(HERE, f) for f in ("chol_oracle.c", "klu_oracle.c")]


def find_openblas():
    try:
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        return os.path.abspath(cands[0]) if cands else None
    except Exception:
        return None


def build(force=False, verbose=True):
    srcs = [s for s in SRCS if os.path.exists(s)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) > os.path.getmtime(s) for s in srcs):
        return LIB
    ob = find_openblas()
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", LIB] + srcs + ["-lm"]
    if ob:
        cmd += [ob, "-Wl,-rpath," + os.path.dirname(ob)]
    else:
        cmd.insert(1, "-DORACLE_NAIVE_BLAS")
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
