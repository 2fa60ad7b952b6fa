"""Builds libtsdslam_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtsdslam_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",               # no FMA contraction: bit parity with the reference's x86-64 build
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-shared", "-cudart", "shared",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(HERE, "..", "include", "tsdslam_b200.h"))
    return d


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("TSD_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ whose wrapper cannot find libgomp; nvcc wants the system g++
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libtsdslam_b200.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
