"""In-tree build of the CUDA library (sm_100a only) and of the oracle (test infrastructure)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpiclas_gpu.so")
SOURCES = ["piclas_gpu.cu", "sort.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc"))) + [os.path.join(ROOT, "include", "piclas_gpu.h")]
STAMP = LIB + ".srchash"   # hash of sources + flags the library was built from (travels with the .so, git-ignored)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # arithmetic contract: no compiler-chosen FMA (DESIGN.md "Arithmetic")
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def source_hash() -> str:
    """sha256 over the compiler flags and every source / header of the library: the build is keyed on content, not on
    modification times (a snapshot copied to another machine keeps the former, not the latter)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [x if os.path.isabs(x) else os.path.join(CSRC, x) for x in HEADERS]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_EXE = os.path.join(HOST_DIR, "run_case")


def build_host(force: bool = False) -> str:
    """The compiled host driver above the C ABI (piclas_b200/host): plain C++17, links libpiclas_gpu.so."""
    deps = [os.path.join(HOST_DIR, f) for f in ("run_case.cpp", "particle_step.hpp", "pgpu_case.hpp", "pgpu_fields.inc")]
    deps += [os.path.join(ROOT, "include", "piclas_gpu.h"), LIB]
    if not force and os.path.exists(HOST_EXE) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_EXE) for d in deps):
        return HOST_EXE
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-o", HOST_EXE, os.path.join(HOST_DIR, "run_case.cpp"), "-L", HERE, "-lpiclas_gpu", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return HOST_EXE


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
    build_oracle()
