"""Build the sm_100a shared library in-tree: compose_b200/libcedr_b200.so.

nvcc cross-compiles without a GPU; the .so is git-ignored but ships to the GPU
box with the gpurun snapshot. -fmad=false is mandatory: the parity contract is
bit-for-bit against the reference built without FMA contraction (SURVEY.md
section 6).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libcedr_b200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("cedr_b200.cu", "tree_plan.cpp")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in
           ("kernels.cuh", "node_solve.cuh", "fast_kernels.cuh", "ring_kernels.cuh", "transposed_kernels.cuh", "cluster_caas.cuh",
            "tree_plan.h")] + \
          [os.path.join(ROOT, "include", f) for f in
           ("cedr_b200.h", "cedr_b200_device_op.h", "cedr_b200_local.hpp")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"),
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS if os.path.exists(f))


def build(force=False, verbose=False, out=None, extra_flags=()):
    """out / extra_flags: a variant build for kernel A/B experiments (CEDR_B200_LIB)."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("CEDR_B200_EXTRA_NVCC_FLAGS", "").split()
    lib = out or LIB
    cmd = [nvcc] + NVCC_FLAGS + extra + list(extra_flags) + \
        (["-Xptxas", "-v"] if verbose else []) + \
        ["-ccbin", "/usr/bin/g++"] * os.path.exists("/usr/bin/g++") + ["-o", lib] + SOURCES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout)
    if r.returncode:
        raise RuntimeError("nvcc failed building %s" % lib)
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
