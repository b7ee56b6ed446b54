"""Build libsrslte_fec_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

    python -m srsran_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared library links the CUDA runtime statically, so it has no
dependency on torch or on a particular libcudart at load time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsrslte_fec_b200.so")
SOURCES = ["engine.cu", "lte_tables.cpp"]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".inc", ".cpp"))) + \
    ["../../include/srslte_b200/batch.h", "../../include/srslte_b200/fec.h"]  # every source and header: a stale library is worse than a slow build
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "-shared", "-cudart", "static"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libsrslte_fec_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
