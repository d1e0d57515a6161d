"""Builds libformoniq_b200.so in-tree with nvcc for sm_100a.

    python -m formoniq_b200.build [--force]

Steps: (1) compile and run the element-tape generator (host C++) to produce
csrc/elmat_gen.cuh, (2) nvcc every .cu to an object, (3) link the shared
library with a static CUDA runtime so it is self-contained next to torch.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libformoniq_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"

SOURCES = ["elmat.cu", "kuhn.cu", "assemble.cu", "tile.cu", "blockop.cu", "matfree.cu", "quadform.cu", "spmv.cu", "blas1.cu", "krylov.cu", "capi.cu"]
HEADERS = ["common.cuh", "internal.hpp", "stream.cuh", "host_widen.hpp", "tape.hpp", "tile_plan.hpp", "kuhn.hpp", "geometry.cuh", "gen_elmat.cpp",
           "../../include/formoniq_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX, "-Xptxas", "-v",
]


def _newer(src: str, dst: str) -> bool:
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return r.stdout


def generate() -> str:
    out = os.path.join(CSRC, "elmat_gen.cuh")
    deps = [os.path.join(CSRC, f) for f in ("gen_elmat.cpp", "tape.hpp")]
    if any(_newer(d, out) for d in deps):
        os.makedirs(OBJDIR, exist_ok=True)
        exe = os.path.join(OBJDIR, "gen_elmat")
        _run([HOST_CXX, "-O1", "-std=c++17", os.path.join(CSRC, "gen_elmat.cpp"), "-o", exe])
        text = _run([exe])
        with open(out, "w") as f:
            f.write(text)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    gen = generate()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [gen]
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        log = _run([NVCC, *NVCC_FLAGS, "-c", src, "-o", obj])
        with open(obj + ".log", "w") as f:
            f.write(log)
        return log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for log in ex.map(compile_one, jobs):
                if verbose:
                    print(log)
    if jobs or force or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-ccbin", HOST_CXX,
              "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
