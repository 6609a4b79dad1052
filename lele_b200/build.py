"""Builds liblele_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python lele_b200/build.py [--force]   (run as a script: importing the package needs the built .so)
nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "liblele_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("LELE_B200_NVCC_DEFS", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "lele_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src: str, force: bool, hdr_m: float) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), hdr_m):
        return obj
    log = obj + ".log"
    with open(log, "w") as lf:
        r = subprocess.run([NVCC, *FLAGS, "-c", srcp, "-o", obj], stdout=lf, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        sys.stderr.write(open(log).read())
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = _deps_mtime()
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr_m), sources()))
    if force or not os.path.exists(SO) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs):
        subprocess.check_call([NVCC, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                               "-cudart", "shared", "-ldl"])
    return SO


if __name__ == "__main__":
    print(build("--force" in sys.argv))
