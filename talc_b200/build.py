"""Builds libtalc_b200.so (hand-written sm_100a kernels + the C ABI of include/talc_b200.h) and the host `talc` CLI.

In-tree, with explicit nvcc flags: the .so travels with the repository snapshot to the GPU box.
-fmad=false: path decisions are made in IEEE double with separate multiply/add, as in the reference's
x86-64 build; the integer kernels do not care.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtalc_b200.so")
CLI = os.path.join(HERE, "_build", "talc")
NVCC = os.environ.get("TALC_NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp"))] + \
           [os.path.join(ROOT, "include", "talc_b200.h")]


def source_hash() -> str:
    """sha256 over the sources the library is compiled from; compiled into the .so (talc_build_source_hash) so that
    a prebuilt binary can be told from one that is stale with respect to the tree it travels with."""
    import hashlib
    h = hashlib.sha256()
    for f in _sources():
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def library_hash(path: str = LIB):
    import ctypes
    try:
        L = ctypes.CDLL(path)
        L.talc_build_source_hash.restype = ctypes.c_char_p
        return L.talc_build_source_hash().decode()
    except Exception:
        return None


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    want = source_hash()
    if force or _stale(LIB, _sources()) or library_hash() != want:
        if not os.path.exists(NVCC):
            # a box without a toolkit may only use a prebuilt library that was compiled from exactly this tree
            if os.path.exists(LIB) and library_hash() == want:
                return LIB
            raise RuntimeError("nvcc not found and libtalc_b200.so is missing or was built from other sources "
                               "(library %s, tree %s)" % (library_hash(), want))
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-DTALC_SOURCE_HASH=\"%s\"" % want, "-o", LIB, os.path.join(CSRC, "talc_b200.cu")]
        subprocess.check_call(cmd, cwd=ROOT)
    return LIB


def build_cli(force: bool = False) -> str:
    src = os.path.join(CSRC, "host", "talc_main.cpp")
    build_library()
    if force or _stale(CLI, [src, os.path.join(CSRC, "host", "reads_io.hpp"), LIB]):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        cmd = [HOSTCXX, "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", CLI,
               "-L", HERE, "-ltalc_b200", "-Wl,-rpath," + HERE]
        subprocess.check_call(cmd, cwd=ROOT)
    return CLI


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_cli(force="--force" in sys.argv)
    print(LIB)
    print(CLI)
