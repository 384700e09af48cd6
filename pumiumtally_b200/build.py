"""In-tree build of libpumitally.so (CUDA kernels + C++ host side + C ABI).

``nvcc`` cross-compiles for sm_100a without a GPU, so this runs in the CPU
container; the resulting ``pumiumtally_b200/lib/libpumitally.so`` travels to
the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")
LIB = os.path.join(LIBDIR, "libpumitally.so")

SOURCES = [
    "walk_kernels.cu",
    "bin_kernels.cu",
    "engine.cu",
    "tet_mesh.cpp",
    "osh_reader.cpp",
    "gmsh_reader.cpp",
    "vtk_writer.cpp",
    "nccl_dl.cpp",
    "c_api.cpp",
    "pumitally_facade.cpp",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fopenmp,-Wall,-fvisibility=default",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpumitally.so cannot be built")


def _newer(a: str, b: str) -> bool:
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit that changed and relink. Returns the .so path."""
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    headers += [os.path.join(ROOT, "include", "pumitally_c.h"),
                os.path.join(ROOT, "include", "pumitally", "PumiTally.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)
    objs, relink = [], force or not os.path.exists(LIB)
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src + ".o")
        objs.append(o)
        if force or _newer(s, o) or newest_header > os.path.getmtime(o):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
            relink = True
    if relink:
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fopenmp", "-lgomp", "-ldl", "-lz",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
