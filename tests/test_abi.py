"""The C-ABI shared library loads without a GPU and exports every symbol include/stp_rasterizer.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import PKG, ROOT

LIB = os.path.join(PKG, "lib", "libstp_rasterizer.so")
HEADER = os.path.join(ROOT, "include", "stp_rasterizer.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|size_t|char\s*\*|void)\s+\*?\s*(stp_[a-z0-9_A-Z]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for n in ("stp_forward", "stp_backward", "stp_mark_visible", "stp_geometry_bytes", "stp_binning_bytes",
              "stp_image_bytes", "stp_view_geometry", "stp_view_binning", "stp_view_image", "stp_requires_cov3D_inv",
              "stp_last_error", "stp_abi_version", "stp_last_timings"):
        assert n in names, n


def test_library_exports_every_declared_symbol(lib):
    for n in declared_functions():
        assert hasattr(lib, n), f"{n} declared in include/stp_rasterizer.h but not exported"


def test_abi_version_and_arena_sizes(lib):
    lib.stp_abi_version.restype = ctypes.c_int
    src = open(HEADER).read()
    assert lib.stp_abi_version() == int(re.search(r"#define STP_ABI_VERSION (\d+)", src).group(1))
    lib.stp_geometry_bytes.restype = ctypes.c_size_t
    lib.stp_geometry_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.stp_image_bytes.restype = ctypes.c_size_t
    lib.stp_image_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.stp_binning_bytes.restype = ctypes.c_size_t
    lib.stp_binning_bytes.argtypes = [ctypes.c_int, ctypes.c_void_p]
    lib.stp_binning_capacity.restype = ctypes.c_int
    lib.stp_binning_capacity.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    P = 1000
    # 87 B per Gaussian of the reference layout minus the 4 B internal_radii and the 4 B point_offsets
    # we do not keep (slots are claimed per tile, binning.cu), +48 B with the inverse covariance
    g0, g1 = lib.stp_geometry_bytes(P, 0), lib.stp_geometry_bytes(P, 1)
    assert g0 >= 79 * P and g1 - g0 >= 48 * P and g1 - g0 < 48 * P + 512
    assert lib.stp_image_bytes(1920, 1080, 0) >= 8 * 1920 * 1080 + 8 * 120 * 68
    # HIER blend log: 8 B per (pixel of a whole tile, record)
    assert lib.stp_image_bytes(1920, 1080, 16) - lib.stp_image_bytes(1920, 1080, 0) >= 120 * 68 * 256 * 16 * 8
    assert lib.stp_binning_bytes(10, None) >= 28 * 10  # point_list + sorted keys + bucket + merge scratch
    assert lib.stp_binning_bytes(0, None) > 0
    # the depth-resorting modes add one 64-byte slab record per instance; capacity <-> size is an exact bijection on
    # multiples of 64 instances (the backward pass re-derives the carve-up from the arena size alone)
    from diff_gaussian_rasterization import _C
    hier = _C.StpSettings(3, *([0] * 12))
    glob = _C.StpSettings(0, *([0] * 12))
    for st in (hier, glob):
        ref = ctypes.addressof(st)
        for cap in (0, 1, 63, 64, 65, 1000, 6381641):
            nbytes = lib.stp_binning_bytes(cap, ref)
            rounded = (cap + 63) // 64 * 64
            assert lib.stp_binning_capacity(nbytes, ref) == rounded
            assert lib.stp_binning_bytes(rounded, ref) == nbytes
    assert (lib.stp_binning_bytes(6400, ctypes.addressof(hier)) - lib.stp_binning_bytes(6400, ctypes.addressof(glob))
            == (64 + 16) * 6400)  # geometric slab record + {r, g, b, id}


def test_settings_validation_without_gpu(lib):
    """unsupported queue sizes are rejected before anything touches the device
    (forward.cu:455-480 / backward.cu:751-760 throw std::runtime_error)."""
    from diff_gaussian_rasterization import _C
    st = _C.StpSettings(3, 0, 64, 9, 4, 0, 0, 0, 0, 0, 0)  # mid queue 9 is not instantiated
    n = ctypes.c_int(0)
    cb = _C.ALLOC_FN(lambda u, b: None)
    rc = _C._lib.stp_forward(cb, None, cb, None, cb, None, 10, 0, 1, None, 16, 16, ctypes.byref(st), None, None, None,
                             None, None, None, 1.0, None, None, None, None, None, None, 1.0, 1.0, 0, None, None, 0, None,
                             ctypes.byref(n))
    assert rc == -2
    assert b"mid queue size" in _C._lib.stp_last_error()
    st = _C.StpSettings(1, 0, 64, 8, 4, 0, 0, 0, 0, 0, 0)  # PPX_FULL backward, backward.cu:733-736
    rc = _C._lib.stp_backward(10, 0, 1, 0, None, 16, 16, ctypes.byref(st), None, *([None] * 5), 1.0, *([None] * 6), 1.0,
                              1.0, *([None] * 15), 0, None)
    assert rc == -2
    assert b"Backward not supported for full per-pixel sort" in _C._lib.stp_last_error()


def test_cpp_interface_is_exported_and_cmake_target_configures(lib, tmp_path):
    """the C++ layer of the reference (CudaRasterizer::Rasterizer, rasterizer.h:184-258) is part of the library, and the
    CMake target of the same name (reference CMakeLists.txt:22-36) configures with the toolchain of this image."""
    import shutil
    import subprocess
    syms = subprocess.run(["nm", "-DC", LIB], capture_output=True, text=True).stdout
    for name in ("CudaRasterizer::Rasterizer::forward(", "CudaRasterizer::Rasterizer::backward(",
                 "CudaRasterizer::Rasterizer::markVisible("):
        assert name in syms, name
    cmake = shutil.which("cmake")
    if cmake is None:
        pytest.skip("cmake not installed")
    out = subprocess.run([cmake, "-S", ROOT, "-B", str(tmp_path / "b"), "-DCMAKE_CUDA_COMPILER=" + (shutil.which("nvcc") or "nvcc")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert (tmp_path / "b" / "Makefile").exists() or (tmp_path / "b" / "build.ninja").exists()


def test_settings_struct_is_the_same_in_header_binding_and_integration_doc():
    """StpSettings is passed by pointer: a binding whose struct is shorter than the header's makes the library read
    garbage (the INTEGRATION.md snippet had gone stale).  Header, ctypes binding and the documented stub
    must list the same fields in the same order."""
    from diff_gaussian_rasterization import _C
    src = open(HEADER).read()
    body = src[src.index("typedef struct StpSettings {"):src.index("} StpSettings;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(int32_t|float)\s+([^;]+);", body):
        for name in decl[1].split(","):
            fields.append((name.strip(), decl[0]))
    binding = [(n, "float" if t is ctypes.c_float else "int32_t") for n, t in _C.StpSettings._fields_]
    assert binding == fields
    assert ctypes.sizeof(_C.StpSettings) == 4 * len(fields)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = doc[doc.index("class StpSettings(ctypes.Structure):"):doc.index("assert lib.stp_abi_version()")]
    doc_fields = re.findall(r'"([a-z_0-9A-Z]+)"', stub)
    assert doc_fields == [n for n, _ in fields]
    assert f"stp_abi_version() == {_C._lib.stp_abi_version()}" in doc


def test_cpp_settings_json_converters_roundtrip(tmp_path):
    """to_json / from_json of CudaRasterizer::SplattingSettings (the reference's rasterizer.h:137-182, used by the viewer to
    load the StopThePop preset files): same keys, missing key -> exception.  Needs some nlohmann/json on the machine."""
    import glob
    import shutil
    import subprocess
    import sysconfig
    cands = glob.glob(os.path.join(sysconfig.get_paths()["purelib"], "include", "**", "nlohmann", "json.hpp"), recursive=True)
    cands += glob.glob("/usr/include/nlohmann/json.hpp")
    if not cands or shutil.which("g++") is None:
        pytest.skip("nlohmann/json.hpp or g++ not available")
    inc = os.path.dirname(os.path.dirname(cands[0]))
    exe = tmp_path / "json_rt"
    subprocess.check_call(["g++", "-std=c++17", "-I" + inc, "-I" + os.path.join(ROOT, "include", "cuda_rasterizer"),
                           os.path.join(ROOT, "tests", "cpp", "settings_json_roundtrip.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    import json
    d = json.loads(out.stdout.strip().splitlines()[0])
    assert d["sort_settings"]["queue_sizes"] == {"per_pixel": 8, "tile_2x2": 8, "tile_4x4": 64}
    assert set(d) == {"sort_settings", "culling_settings", "load_balancing", "proper_ewa_scaling"}
