#!/usr/bin/env python
"""Build libstp_rasterizer.so (sm_100a) in-tree with plain nvcc; no torch headers involved.

    python stopthepop-rasterization_b200/csrc/build.py [--force] [--verbose]

Outputs: stopthepop-rasterization_b200/lib/libstp_rasterizer.so (+ obj/*.o), both git-ignored but
shipped to the GPU box by gpurun.
"""
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libstp_rasterizer.so")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "render_global.cu", "render_hier.cu", "render_ppx.cu",
           "preprocess_bwd.cu", "debug_vis.cu", "rasterizer_shim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--compiler-options", "-fPIC", "-Xptxas", "-v", "-Xcudafe", "--diag_suppress=177"]


def newest_header():
    return max(os.path.getmtime(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith((".cuh", ".h")))


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    inc_time = max(newest_header(), os.path.getmtime(os.path.join(PKG, "..", "include", "stp_rasterizer.h")))
    logs = {}

    def compile_one(src):
        path = os.path.join(HERE, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), inc_time):
            return obj, False
        t0 = time.time()
        p = subprocess.run(["nvcc", "-c", path, "-o", obj] + NVCC_FLAGS, capture_output=True, text=True)
        logs[src] = p.stderr
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            print(f"[build] {src}: {time.time() - t0:.1f}s")
        return obj, True

    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(LIB):
        subprocess.check_call(["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                         "-lcudart"])
    with open(os.path.join(LIB_DIR, "ptxas.log"), "a" if not force else "w") as fh:
        for k, v in logs.items():
            fh.write(f"==== {k}\n{v}\n")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose=True)
    print(lib)
