"""Builds redsec_b200/libredsec_b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

Called by __graft_entry__.build(); cross-compiles without a GPU.  The .so is git-ignored but travels
to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libredsec_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_SOURCES = ["api.cu"]
CXX_SOURCES = ["client.cpp", "layers.cpp", "comm.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp,-O3", "-shared",
]


def _sources():
    out = []
    for f in CUDA_SOURCES + CXX_SOURCES:
        p = os.path.join(CSRC, f)
        if os.path.exists(p):
            out.append(p)
    return out


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "redsec_b200.h"))
    return deps


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    extra = os.environ.get("RS_NVCC_EXTRA", "").split()      # e.g. -DRS_WAIT_HINT_NS=200u for tuning experiments
    cmd = [NVCC] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + _sources() + ["-lgomp", "-ldl"]
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc/bin/* (a wrapper without OpenMP specs); use the system g++ as nvcc's host compiler
    cmd[1:1] = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
