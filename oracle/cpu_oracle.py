"""numpy front-end of the plain-C restatement oracle (oracle/stp_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libstp_oracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


class OrcSettings(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("sort_mode", "sort_order", "q_tile4", "q_mid", "q_head", "rect_bounding",
                                            "tight_opacity_bounding", "tile_based_culling", "hier_culling",
                                            "load_balancing", "proper_ewa_scaling")]


_fp = ctypes.POINTER(ctypes.c_float)


class OrcInputs(ctypes.Structure):
    _fields_ = [("P", ctypes.c_int), ("D", ctypes.c_int), ("M", ctypes.c_int), ("W", ctypes.c_int), ("H", ctypes.c_int),
                ("means3D", _fp), ("scales", _fp), ("rotations", _fp), ("opacities", _fp), ("shs", _fp),
                ("colors_precomp", _fp), ("cov3D_precomp", _fp), ("scale_modifier", ctypes.c_float),
                ("viewmatrix", _fp), ("projmatrix", _fp), ("inv_viewproj", _fp), ("campos", _fp), ("bg", _fp),
                ("tan_fovx", ctypes.c_float), ("tan_fovy", ctypes.c_float)]


class OrcState(ctypes.Structure):
    _fields_ = [("P", ctypes.c_int), ("R", ctypes.c_int), ("tiles", ctypes.c_int), ("W", ctypes.c_int), ("H", ctypes.c_int),
                ("radii", ctypes.POINTER(ctypes.c_int)),
                ("depths", _fp), ("means2D", _fp), ("rects2D", _fp), ("conic_opacity", _fp), ("rgb", _fp), ("cov3D", _fp),
                ("cov3D_inv", _fp), ("clamped", ctypes.POINTER(ctypes.c_uint8)),
                ("tiles_touched", ctypes.POINTER(ctypes.c_uint32)), ("point_offsets", ctypes.POINTER(ctypes.c_uint32)),
                ("keys", ctypes.POINTER(ctypes.c_uint64)), ("point_list", ctypes.POINTER(ctypes.c_uint32)),
                ("ranges", ctypes.POINTER(ctypes.c_uint32)), ("out_color", _fp), ("final_T", _fp),
                ("n_contrib", ctypes.POINTER(ctypes.c_uint32))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "stp_oracle.c")):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_forward.restype = ctypes.POINTER(OrcState)
        _lib.orc_forward.argtypes = [ctypes.POINTER(OrcInputs), ctypes.POINTER(OrcSettings)]
        _lib.orc_backward.restype = ctypes.c_int
        _lib.orc_free.argtypes = [ctypes.POINTER(OrcState)]
        _lib.orc_debug_visualisation.restype = ctypes.POINTER(OrcState)
        _lib.orc_debug_visualisation.argtypes = [ctypes.POINTER(OrcInputs), ctypes.POINTER(OrcSettings), ctypes.c_int, _fp]
    return _lib


def settings_struct(d):
    ss, cs, q = d["sort_settings"], d["culling_settings"], d["sort_settings"]["queue_sizes"]
    return OrcSettings(int(ss["sort_mode"]), int(ss["sort_order"]), int(q["tile_4x4"]), int(q["tile_2x2"]),
                       int(q["per_pixel"]), int(cs["rect_bounding"]), int(cs["tight_opacity_bounding"]),
                       int(cs["tile_based_culling"]), int(cs["hierarchical_4x4_culling"]), int(d["load_balancing"]),
                       int(d["proper_ewa_scaling"]))


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


# DebugVisualization codes of include/stp_rasterizer.h (STP_DEBUG_*)
SORT_ERROR_OPACITY, SORT_ERROR_DISTANCE, COUNT_PER_TILE, DEPTH, COUNT_PER_PIXEL, TRANSMITTANCE = 1, 2, 3, 4, 5, 6

# Magma: degree-6 polynomial per channel, M. Zucker's public-domain fit of matplotlib's table -- the coefficients the
# reference evaluates in colormapMagma (stopthepop_common.cuh:623-642)
_MAGMA = np.array([[-0.002136485053939582, -0.000749655052795221, -0.005386127855323933],
                   [0.2516605407371642, 0.6775232436837668, 2.494026599312351],
                   [8.353717279216625, -3.577719514958484, 0.3144679030132573],
                   [-27.66873308576866, 14.26473078096533, -13.64921318813922],
                   [52.17613981234068, -27.94360607168351, 12.94416944238394],
                   [-50.76852536473588, 29.04658282127291, 4.23415299384598],
                   [18.65570506591883, -11.48977351997711, -5.601961508734096]], dtype=np.float32)
_turbo = None


def turbo_table():
    """Google's published 256-entry Turbo sRGB table (A. Mikhailov 2019, Apache-2.0), which the reference embeds in
    colormapTurbo (stopthepop_common.cuh:645-656).  Read from the one copy of the table in this repository; the golden
    render_depth images of the reference build (tests/golden/depth_vis_*.npz) pin it."""
    global _turbo
    if _turbo is None:
        import re
        path = os.path.join(_HERE, "..", "stopthepop-rasterization_b200", "csrc", "stp_turbo_lut.cuh")
        body = open(path).read().split("kTurboLut[256][3]", 1)[1]
        vals = [float(x) for x in re.findall(r"(\d\.\d+)f", body)]
        assert len(vals) == 768
        _turbo = np.array(vals, dtype=np.float32).reshape(256, 3)
    return _turbo


def colormap(value, T, kind, debug_range=None):
    """render_debug_CUDA, forward.cu:674-714: value / T [H,W] float32 -> [3,H,W].  Depth: Turbo of
    clamp(value + T * max, min, max) / (max - min); the others: Magma of clamp(value, min, max) / (max - min), with
    (min, max) of the value plane (applyDebugVisualization, rasterizer_impl.cu:66-76) or the caller's (debug_normalize)."""
    f = np.float32
    mn, mx = (f(value.min()), f(value.max())) if debug_range is None else (f(debug_range[0]), f(debug_range[1]))
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == DEPTH:
            x = np.minimum(np.maximum(value + T * mx, mn), mx) / f(mx - mn)
            interp = np.minimum(np.maximum(x * f(255.0), f(0.0)), f(255.0))
            lo = np.where(x > 0, interp.astype(np.int32), 0)
            hi = np.where(lo >= 255, 255, lo + 1)
            diff = (interp - lo.astype(f))[..., None]
            lut = turbo_table()
            rgb = np.clip(lut[lo] + (lut[hi] - lut[lo]) * diff, f(0), f(1))
        else:
            x = np.clip(np.minimum(np.maximum(value, mn), mx) / f(mx - mn), f(0), f(1))[..., None]
            rgb = np.broadcast_to(_MAGMA[6], x.shape[:-1] + (3,)).astype(f)
            for k in range(5, -1, -1):
                rgb = _MAGMA[k] + x * rgb
            rgb = np.clip(rgb, f(0), f(1))
    return np.ascontiguousarray(np.moveaxis(rgb.astype(f), -1, 0))


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


class Oracle:
    """one forward evaluation; keeps the C state alive so backward() can be called on it."""

    def __init__(self, settings, means3D, scales, rotations, opacities, shs, sh_degree, viewmatrix, projmatrix,
                 inv_viewprojmatrix, campos, bg, tanfovx, tanfovy, W, H, colors_precomp=None, cov3D_precomp=None,
                 scale_modifier=1.0):
        self.keep = [_c(x) for x in (means3D, scales, rotations, opacities, shs, colors_precomp, cov3D_precomp, viewmatrix,
                                     projmatrix, inv_viewprojmatrix, campos, bg)]
        m3, sc, ro, op, sh, cp, cv, vm, pm, iv, cam, bgc = self.keep
        self.P = m3.shape[0]
        self.M = 0 if sh is None else sh.shape[1]
        self.W, self.H = int(W), int(H)
        self.inp = OrcInputs(self.P, int(sh_degree), self.M, self.W, self.H, _p(m3), _p(sc), _p(ro), _p(op), _p(sh), _p(cp),
                             _p(cv), float(scale_modifier), _p(vm), _p(pm), _p(iv), _p(cam), _p(bgc), float(tanfovx),
                             float(tanfovy))
        self.settings = settings_struct(settings)
        self.st = lib().orc_forward(ctypes.byref(self.inp), ctypes.byref(self.settings))
        s = self.st.contents
        P, R, N, T = self.P, s.R, self.W * self.H, s.tiles
        arr = lambda ptr, n, dt: np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dt).copy()  # noqa: E731
        self.R = R
        self.radii = arr(s.radii, P, np.int32)
        self.depths = arr(s.depths, P, np.float32)
        self.means2D = arr(s.means2D, 2 * P, np.float32).reshape(P, 2)
        self.rects2D = arr(s.rects2D, 2 * P, np.float32).reshape(P, 2)
        self.conic_opacity = arr(s.conic_opacity, 4 * P, np.float32).reshape(P, 4)
        self.rgb = arr(s.rgb, 3 * P, np.float32).reshape(P, 3)
        self.clamped = arr(s.clamped, 3 * P, np.uint8).reshape(P, 3)
        self.tiles_touched = arr(s.tiles_touched, P, np.int64).astype(np.int32)
        self.point_list = arr(s.point_list, R, np.int64).astype(np.uint32).view(np.int32)
        self.keys = arr(s.keys, R, np.uint64).view(np.int64)
        self.ranges = arr(s.ranges, 2 * T, np.int64).astype(np.int32).reshape(T, 2)
        self.out_color = arr(s.out_color, 3 * N, np.float32).reshape(3, self.H, self.W)
        self.final_T = arr(s.final_T, N, np.float32).reshape(self.H, self.W)
        self.n_contrib = arr(s.n_contrib, N, np.int64).astype(np.int32).reshape(self.H, self.W)

    def backward(self, dL_dout, pixel_colors=None, full_sort_ext=False):
        """full_sort_ext=True: PPX_FULL backward, which the reference does NOT have (backward.cu:733-736) -- the derived
        extension of stp_oracle.c:render_full (same per-pixel order as the forward pass, the reference's front-to-back
        gradient terms), used to pin the CUDA path's replay backward."""
        P, M = self.P, self.M
        lib().orc_allow_full_backward_ext(1 if full_sort_ext else 0)
        pc = _c(self.out_color if pixel_colors is None else pixel_colors)
        dl = _c(dL_dout)
        z = lambda *s: np.zeros(s, dtype=np.float32)  # noqa: E731
        g = dict(dL_dmeans2D=z(P, 3), dL_dconic=z(P, 4), dL_dopacity=z(P, 1), dL_dcolors=z(P, 3), dL_dmeans3D=z(P, 3),
                 dL_dcov3D=z(P, 6), dL_dsh=z(P, max(M, 1), 3), dL_dscales=z(P, 3), dL_drot=z(P, 4))
        rc = lib().orc_backward(ctypes.byref(self.inp), ctypes.byref(self.settings), self.st, _p(pc), _p(dl),
                                _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
                                _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]),
                                _p(g["dL_drot"]))
        if rc != 0:
            raise RuntimeError("Backward not supported for full per-pixel sort")
        g["dL_dsh"] = g["dL_dsh"][:, :M]
        return g

    def debug_visualisation(self, kind, debug_range=None):
        """The ENABLE_DEBUG_VIZ forward pass (accumSortingErrorDepth / outputDebugVis, stopthepop_common.cuh:264-307) of
        this scene: returns dict(value[H,W], T[H,W], stats=(min, max, mean, std) of value as the viewer's callback gets
        them (rasterizer_impl.cu:66-97), image[3,H,W] = the colour-mapped frame)."""
        raw = np.zeros((2, self.H * self.W), dtype=np.float32)
        st = lib().orc_debug_visualisation(ctypes.byref(self.inp), ctypes.byref(self.settings), int(kind), _p(raw))
        lib().orc_free(st)
        value, T = raw[0].reshape(self.H, self.W), raw[1].reshape(self.H, self.W)
        v64 = value.astype(np.float64)
        stats = (float(value.min()), float(value.max()), float(v64.mean()), float(v64.std()))
        return dict(value=value, T=T, stats=stats, image=colormap(value, T, int(kind), debug_range))

    def close(self):
        if self.st:
            lib().orc_free(self.st)
            self.st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
