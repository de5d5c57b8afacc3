"""Builds libp2e.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

The built library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libp2e.so")
SOURCES = ["p2e_kernels.cu", "p2e_chain.cu", "p2e_ux.cu", "p2e_ntt.cu", "p2e_genstream.cu", "p2e_tape.cpp"]
HEADERS = ["p2e_field.cuh", "p2e_f8.cuh", "p2e_f8inv.cuh", "p2e_fold.cuh", "p2e_common.h", "p2e_tape.h", os.path.join("..", "..", "include", "p2e.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the package into plonky2_ecdsa_b200/libp2e.so."""
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(CSRC, os.path.splitext(s)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    # shared cudart: the artefact carries only the runtime symbols it imports (and shares the runtime instance of a
    # torch host process); the rpath lets a plain C host (examples/host_demo.c) find it without torch
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    cmd = [nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                    "-Xlinker", "-rpath=" + cuda_lib]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
