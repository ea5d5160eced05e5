"""In-tree build of libnem_b200.so (CUDA kernels for sm_100a + C host) and the helper binaries.

The library is built with explicit nvcc/gcc commands (no JIT cache) so that the .so travels to
the GPU box with the repository snapshot.  ``python -m pangenomenem_b200.build`` rebuilds.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libnem_b200.so")
CLI = os.path.join(PKG, "nem_exe")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
CC_FLAGS = ["-O2", "-std=gnu11", "-Wall", "-Wno-unused-function", "-fPIC", "-ffp-contract=off",
            "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", os.path.join(CUDA_HOME, "include")]

CU_SOURCES = ["nem_kernels.cu", "nem_sub_kernels.cu"]
C_SOURCES = ["nem_fit.c", "nem_resample.c", "nem_comm.c", "nem_io.c", "nem_api.c"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_all(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(ROOT, "include", "nem_b200.h")] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    objs = []
    for f in CU_SOURCES:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        if force or _newer([src] + headers, obj):
            _run([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj], verbose)
        objs.append(obj)
    for f in C_SOURCES:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        if not os.path.exists(src):
            continue
        if force or _newer([src] + headers, obj):
            _run(["gcc"] + CC_FLAGS + ["-c", src, "-o", obj], verbose)
        objs.append(obj)
    if force or _newer(objs, LIB):
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
             + ["-lm", "-ldl", "-lpthread"], verbose)
    cli_src = os.path.join(CSRC, "nem_cli.c")
    if os.path.exists(cli_src) and (force or _newer([cli_src, LIB], CLI)):
        _run(["gcc"] + CC_FLAGS + [cli_src, "-o", CLI, "-L", PKG, "-lnem_b200",
                                   "-Wl,-rpath,$ORIGIN", "-lm"], verbose)
    return LIB


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
