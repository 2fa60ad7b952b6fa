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


STAMP = LIB + ".srchash"


def source_hash() -> str:
    """Hash of everything the library is built from (contents, not mtimes: a snapshot copied to another box does
    not keep the order of modification times)."""
    import hashlib
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + os.environ.get("TSD_NVCC_EXTRA", "").split()).encode())
    for p in sorted(_deps()):
        if os.path.isfile(p):
            h.update(os.path.basename(p).encode())
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    try:
        with open(STAMP) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds under an exclusive file lock and moves the result into place atomically: several ranks importing the
    package at once (torchrun) either wait for the one build or find it done."""
    import fcntl
    if not force and not needs_build():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # somebody else built it while this process waited
                return LIB
            nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
            tmp = LIB + f".tmp{os.getpid()}"
            cmd = ([nvcc] + NVCC_FLAGS + os.environ.get("TSD_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) +
                   ["-o", tmp] + sources())
            env = dict(os.environ)
            # the image exports CXX=/opt/gcc/bin/g++ whose wrapper cannot find libgomp; nvcc wants the system g++
            env.pop("CXX", None)
            env.pop("CC", None)
            # host compiler: TSD_CCBIN if set, else the system g++ where it exists, else nvcc's own default
            ccbin = os.environ.get("TSD_CCBIN") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None)
            r = subprocess.run(cmd + (["-ccbin", ccbin] if ccbin else []), capture_output=True, text=True, env=env)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libtsdslam_b200.so")
            os.replace(tmp, LIB)
            with open(STAMP + ".tmp", "w") as f:
                f.write(source_hash())
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
