#!/usr/bin/env python
"""Build the UNMODIFIED reference CUDA extension into oracle/_ref/ (test infrastructure only).

The reference (r4dl/StopThePop-Rasterization) ships no CPU path and no tests, so the numeric
oracle for this repo is the reference's own CUDA build.  This script
compiles the five translation units the reference's setup.py lists (setup.py:24-29), *where
they lie* under /root/reference, with plain nvcc/g++ command lines that mirror what
torch.utils.cpp_extension.CUDAExtension would pass (no reference build system is run, no
reference source is copied into this repository).  Outputs go only to oracle/_ref/:

    oracle/_ref/_C.cpython-312-x86_64-linux-gnu.so     the reference pybind module
    oracle/_ref/obj/*.o                                intermediate objects
    oracle/_ref/BUILD_INFO.json                        flags + source commit, for DESIGN.md

oracle/_ref/ is git-ignored (it is not product source) but NOT gpurun-ignored, so the built
module travels to the GPU box, where /root/reference does not exist.

Only tests/, __graft_entry__.smoke() and bench.py (--impl reference / cpu_baseline) may load it.
"""
import json
import os
import subprocess
import sys
import sysconfig
import time
from concurrent.futures import ThreadPoolExecutor

REF = os.environ.get("STP_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")

SOURCES = [  # setup.py:24-29 of the reference
    "cuda_rasterizer/rasterizer_impl.cu",
    "cuda_rasterizer/forward.cu",
    "cuda_rasterizer/backward.cu",
    "rasterize_points.cu",
    "ext.cpp",
]


def main() -> int:
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present; keeping whatever is prebuilt in {OUT}")
        return 0
    import torch  # noqa: F401  (needed for include paths / ABI flag)
    from torch.utils.cpp_extension import COMMON_NVCC_FLAGS, include_paths, library_paths

    os.makedirs(OBJ, exist_ok=True)
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(OUT, "_C" + ext_suffix)
    srcs = [os.path.join(REF, s) for s in SOURCES]
    if os.path.exists(target) and all(os.path.getmtime(target) > os.path.getmtime(s) for s in srcs):
        print(f"[build_ref] up to date: {target}")
        return 0

    incs = [f"-I{p}" for p in include_paths("cuda")]
    incs += [f"-I{sysconfig.get_paths()['include']}",
             f"-I{os.path.join(REF, 'third_party')}",
             f"-I{os.path.join(REF, 'third_party/glm')}",
             f"-I{REF}"]
    defs = ["-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    # the reference passes no -gencode; torch would use TORCH_CUDA_ARCH_LIST.  sm_100 = plain
    # (non-"a") Blackwell target: this is "the reference recompiled for the box", nothing more.
    arch = ["-gencode", "arch=compute_100,code=sm_100"]
    nvcc_flags = COMMON_NVCC_FLAGS + ["--compiler-options", "-fPIC", "-std=c++17", "-O3", "-w"] + arch
    cxx_flags = ["-fPIC", "-std=c++17", "-O2", "-w"]

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src).replace(".", "_") + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src):
            return obj
        if src.endswith(".cu"):
            cmd = ["nvcc", "-c", src, "-o", obj] + incs + defs + nvcc_flags
        else:
            cmd = ["g++", "-c", src, "-o", obj] + incs + defs + cxx_flags
        t0 = time.time()
        subprocess.check_call(cmd)
        print(f"[build_ref] {os.path.basename(src)}: {time.time() - t0:.0f}s", flush=True)
        return obj

    t0 = time.time()
    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(compile_one, srcs))
    libs = [f"-L{p}" for p in library_paths("cuda")]
    link = ["g++", "-shared", "-o", target] + objs + libs + [
        "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    subprocess.check_call(link)
    commit = None
    try:
        commit = json.load(open(os.path.join(REF, ".SUBMODULES.json")))
    except Exception:
        pass
    json.dump({"target": os.path.basename(target), "nvcc_flags": nvcc_flags, "cxx_flags": cxx_flags,
               "sources": SOURCES, "reference": commit, "seconds": round(time.time() - t0, 1)},
              open(os.path.join(OUT, "BUILD_INFO.json"), "w"), indent=1)
    print(f"[build_ref] built {target} in {time.time() - t0:.0f}s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
