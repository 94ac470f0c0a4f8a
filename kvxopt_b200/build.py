"""Build libb200sparse.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by hand:
    python -m kvxopt_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200sparse.so")
SOURCES = ["ordering.cpp", "symbolic.cpp", "klu_host.cpp", "chol_gpu.cu", "klu_gpu.cu", "capi.cu", "klu_capi.cu", "kkt_gpu.cu", "kktd_gpu.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC,-O3,", "--expt-relaxed-constexpr"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".hpp", ".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "b200sparse.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((s, subprocess.Popen(cmd)))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on " + s)
    if force or procs or _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
