#!/usr/bin/env python
"""Build an A/B variant of the library: recompile the named source files with extra -D flags and link them with the
standard objects -> stopthepop-rasterization_b200/lib/var/libstp_<name>.so  (used with STP_RASTERIZER_LIB, tools/gpu_variants.sh)
usage: python tools/build_variant.py <name> <file.cu[,file2.cu]> [-DFOO=1 ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "stopthepop-rasterization_b200", "csrc")
LIB = os.path.join(ROOT, "stopthepop-rasterization_b200", "lib")
sys.path.insert(0, CS)
import build as B  # noqa: E402
name, files, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
alt = {f.split("=")[0]: f.split("=")[1] for f in files if "=" in f}  # file.cu=/path/to/alternative/source.cu
files = [f.split("=")[0] for f in files]
if not os.environ.get('STP_VARIANT_NO_BASE'):
    B.build()
vdir = os.path.join(LIB, "var")
os.makedirs(os.path.join(vdir, "obj_" + name), exist_ok=True)
objs = []
for src in B.SOURCES:
    if src in files:
        o = os.path.join(vdir, "obj_" + name, src.replace(".cu", ".o"))
        subprocess.check_call(["nvcc", "-c", alt.get(src, os.path.join(CS, src)), "-I", CS, "-o", o] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags)
        objs.append(o)
    else:
        objs.append(os.path.join(B.OBJ_DIR, src.replace(".cu", ".o")))
out = os.path.join(vdir, f"libstp_{name}.so")
subprocess.check_call(["nvcc", "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
print(out)
