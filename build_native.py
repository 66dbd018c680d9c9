"""Builds pressurepoissonsolver_b200/libtgpu.so (the C-ABI library, include/tgpu.h) for sm_100a.

nvcc cross-compiles without a GPU.  The .so is kept in-tree (git-ignored) so it travels to the GPU
box with the repo snapshot.  Usage: python build_native.py [--force] [-v]
(kept outside the package so that building never needs the library to be importable)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, "pressurepoissonsolver_b200")
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtgpu.so")
SOURCES = [os.path.join(CSRC, "tgpu.cu"), os.path.join(CSRC, "mesh.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, n) for n in sorted(os.listdir(CSRC)) if n.endswith((".cuh", ".h"))] + [
    os.path.join(ROOT, "include", "tgpu.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-cudart", "static"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building libtgpu.so")
    if verbose:
        print(res.stdout)
    return LIB


APP = os.path.join(ROOT, "apps", "steady")


def build_app(force=False):
    """C++ host program against the mirrored plugin surface (include/tgpu_plugin.hpp)."""
    src = os.path.join(ROOT, "apps", "steady.cpp")
    deps = [src, os.path.join(ROOT, "include", "tgpu_plugin.hpp"), os.path.join(ROOT, "include", "tgpu.h"), LIB]
    if not force and os.path.exists(APP) and all(os.path.getmtime(d) <= os.path.getmtime(APP) for d in deps):
        return APP
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++14", "-I", os.path.join(ROOT, "include"), src, "-o", APP,
           "-L", HERE, "-ltgpu", "-Wl,-rpath," + HERE]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("g++ failed building apps/steady")
    return APP


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_app(force="--force" in sys.argv)
    print(LIB)
