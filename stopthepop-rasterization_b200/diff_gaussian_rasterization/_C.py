"""ctypes binding of libstp_rasterizer.so with the call surface of the reference's pybind module.

Mirrors ``diff_gaussian_rasterization._C`` of r4dl/StopThePop-Rasterization (ext.cpp:15-19):

    rasterize_gaussians(...22 positional args...)  -> (num_rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer)
    rasterize_gaussians_backward(...25 args...)    -> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
    mark_visible(means3D, viewmatrix, projmatrix)  -> bool[P]

(signatures: rasterize_points.h:26-82; tensor allocation and marshalling: rasterize_points.cu:43-253).
PyTorch is used for device memory and the current stream only; all compute happens in the hand
written sm_100a kernels behind the C ABI declared in include/stp_rasterizer.h.  There is no CPU or
eager fallback: a missing library is an ImportError.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("STP_RASTERIZER_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libstp_rasterizer.so"))

if not os.path.exists(_LIB_PATH):
    raise ImportError(
        f"{_LIB_PATH} not found: build the CUDA library first "
        "(python -c 'import __graft_entry__ as g; g.build()' or python stopthepop-rasterization_b200/csrc/build.py). "
        "There is no CPU fallback for the rasterizer.")

_lib = ctypes.CDLL(_LIB_PATH)

NUM_CHANNELS = 3  # config.h:15


class StpSettings(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "sort_mode", "sort_order", "queue_tile_4x4", "queue_tile_2x2", "queue_per_pixel", "rect_bounding",
        "tight_opacity_bounding", "tile_based_culling", "hierarchical_4x4_culling", "load_balancing",
        "proper_ewa_scaling", "blend_record_cap", "debug_visualization", "debug_normalize")] + [
        ("debug_min", ctypes.c_float), ("debug_max", ctypes.c_float), ("debug_pixel_x", ctypes.c_int32),
        ("debug_pixel_y", ctypes.c_int32)]


class StpTileBand(ctypes.Structure):
    _fields_ = [("row_begin", ctypes.c_int32), ("row_end", ctypes.c_int32)]


class StpGeometryView(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "depths", "clamped", "rects2D", "means2D", "cov3D", "cov3D_inv", "conic_opacity", "rgb", "tiles_touched")]


class StpBinningView(ctypes.Structure):
    _fields_ = [("point_list", ctypes.c_void_p), ("point_list_keys", ctypes.c_void_p)]


class StpImageView(ctypes.Structure):
    _fields_ = [("final_T", ctypes.c_void_p), ("n_contrib", ctypes.c_void_p), ("ranges", ctypes.c_void_p)]


ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

_P = ctypes.c_void_p
_lib.stp_last_error.restype = ctypes.c_char_p
_lib.stp_abi_version.restype = ctypes.c_int
_lib.stp_geometry_bytes.restype = ctypes.c_size_t
_lib.stp_geometry_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
_lib.stp_binning_bytes.restype = ctypes.c_size_t
_lib.stp_binning_bytes.argtypes = [ctypes.c_int, ctypes.POINTER(StpSettings)]
_lib.stp_binning_capacity.restype = ctypes.c_int
_lib.stp_binning_capacity.argtypes = [ctypes.c_size_t, ctypes.POINTER(StpSettings)]
_lib.stp_image_bytes.restype = ctypes.c_size_t
_lib.stp_image_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
_lib.stp_requires_cov3D_inv.argtypes = [ctypes.POINTER(StpSettings)]
_lib.stp_note_num_rendered.restype = None
_lib.stp_note_num_rendered.argtypes = [ctypes.c_int]
_lib.stp_set_num_rendered_hint.restype = None
_lib.stp_set_num_rendered_hint.argtypes = [ctypes.c_int]
_lib.stp_view_geometry.argtypes = [_P, ctypes.c_int, ctypes.c_int, ctypes.POINTER(StpGeometryView)]
_lib.stp_view_binning.argtypes = [_P, ctypes.c_int, ctypes.POINTER(StpBinningView)]
_lib.stp_view_image.argtypes = [_P, ctypes.c_int, ctypes.c_int, ctypes.POINTER(StpImageView)]
_lib.stp_mark_visible.argtypes = [ctypes.c_int, _P, _P, _P, _P, _P]
_lib.stp_last_timings.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_char_p), ctypes.c_int]
_lib.stp_timing_summary.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_char_p),
                                    ctypes.POINTER(ctypes.c_int), ctypes.c_int]
_lib.stp_timing_reset.restype = None
_lib.stp_kernel_launches.restype = ctypes.c_longlong
_lib.stp_forward.restype = ctypes.c_int
_lib.stp_forward.argtypes = [
    ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P,  # arenas
    ctypes.c_int, ctypes.c_int, ctypes.c_int,  # P D M
    _P, ctypes.c_int, ctypes.c_int,            # background W H
    ctypes.POINTER(StpSettings), ctypes.POINTER(StpTileBand),
    _P, _P, _P, _P, _P, ctypes.c_float, _P, _P,  # means3D shs colors opac scales mod rot cov3D
    _P, _P, _P, _P, ctypes.c_float, ctypes.c_float, ctypes.c_int,  # view proj inv campos tanx tany prefiltered
    _P, _P, ctypes.c_int, _P, _P]  # out_color radii debug stream num_rendered (host int* / pinned int[2])
_lib.stp_backward.restype = ctypes.c_int
_lib.stp_backward.argtypes = [
    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,  # P D M binning_bytes
    _P, ctypes.c_int, ctypes.c_int, ctypes.POINTER(StpSettings), ctypes.POINTER(StpTileBand),
    _P, _P, _P, _P, _P, ctypes.c_float, _P, _P,  # means3D shs opac colors scales mod rot cov3D
    _P, _P, _P, _P, ctypes.c_float, ctypes.c_float,  # view proj inv campos tanx tany
    _P, _P, _P, _P, _P, _P,  # pixel_colors radii geom binning image dL_dpix
    _P, _P, _P, _P, _P, _P, _P, _P, _P,  # 9 grads
    ctypes.c_int, _P]

_lib.stp_backward_render.restype = ctypes.c_int
_lib.stp_backward_render.argtypes = list(_lib.stp_backward.argtypes)
_lib.stp_backward_preprocess.restype = ctypes.c_int
_lib.stp_backward_preprocess.argtypes = list(_lib.stp_backward.argtypes) + [ctypes.c_int, ctypes.c_int]

if _lib.stp_abi_version() != 7:
    raise ImportError("libstp_rasterizer.so ABI version mismatch")

LIBRARY_PATH = _LIB_PATH


def _err():
    return _lib.stp_last_error().decode("utf-8", "replace")


# Blend log: blends recorded per pixel by a forward pass that will be followed by a backward pass (8 B each; the image
# arena grows by 2 KB per pixel at the default).  Pixels that blend more fall back to the list-driven backward kernels.
# HIER: on by default (backward 4.7x faster at C3b: no second hierarchical re-sort).  GLOBAL: implemented and tested but
# off by default -- measured slower on B200 (C2: fwd 0.35->0.47 ms, bwd 0.96->1.06 ms), because the list-driven GLOBAL
# backward already pre-reduces every Gaussian's gradient across the warp before touching memory.
BLEND_RECORD_CAP = int(os.environ.get("STP_BLEND_RECORD_CAP", "256"))
# PPX_FULL: on -- the log is what makes a backward pass possible at all (the reference has none, backward.cu:733-736).
BLEND_RECORD_MODES = (0, 1, 3) if os.environ.get("STP_BLEND_RECORD_GLOBAL", "0") == "1" else (1, 3)


# DebugVisualization (rasterizer_debug.h:11-20) -> STP_DEBUG_* of include/stp_rasterizer.h
STP_DEBUG_SORT_ERROR_OPACITY, STP_DEBUG_SORT_ERROR_DISTANCE, STP_DEBUG_COUNT_PER_TILE = 1, 2, 3
STP_DEBUG_DEPTH, STP_DEBUG_COUNT_PER_PIXEL, STP_DEBUG_TRANSMITTANCE = 4, 5, 6
_VIS_WITHOUT_LOG = (STP_DEBUG_COUNT_PER_TILE, STP_DEBUG_TRANSMITTANCE)


def settings_from_dict(d, blend_record_cap=0, render_depth=False, debug_visualization=0, debug_range=None):
    """dict produced by ExtendedSettings.to_dict() -> StpSettings; every key mandatory like the
    reference's from_json (rasterizer.h:160-182 uses .at()).  render_depth=True is DebugVisualization::Depth
    (rasterize_points.cu:104-107); the other visualisation types are only reachable through the C / C++ interface in the
    reference -- here also through the private `debug_visualization` argument of rasterize_gaussians."""
    ss, cs = d["sort_settings"], d["culling_settings"]
    q = ss["queue_sizes"]
    vis = STP_DEBUG_DEPTH if render_depth else int(debug_visualization or 0)
    keep_log = int(ss["sort_mode"]) in BLEND_RECORD_MODES or (vis != 0 and vis not in _VIS_WITHOUT_LOG)
    lo, hi = debug_range if debug_range is not None else (0.0, 10000.0)
    return StpSettings(int(ss["sort_mode"]), int(ss["sort_order"]), int(q["tile_4x4"]), int(q["tile_2x2"]),
                       int(q["per_pixel"]), int(bool(cs["rect_bounding"])), int(bool(cs["tight_opacity_bounding"])),
                       int(bool(cs["tile_based_culling"])), int(bool(cs["hierarchical_4x4_culling"])),
                       int(bool(d["load_balancing"])), int(bool(d["proper_ewa_scaling"])),
                       int(blend_record_cap) if keep_log else 0, vis, int(debug_range is not None), float(lo), float(hi),
                       0, 0)


def _ptr(t):
    """device pointer of a tensor, or NULL for the reference's 'absent' convention (empty tensor)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32(t, device):
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError("expected a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError("expected a float32 tensor")
    return t.contiguous()


class _Arena:
    """one resizable byte buffer handed to the library through the allocation callback
    (resizeFunctional, rasterize_points.cu:33-41)."""
    __slots__ = ("device", "tensor", "key")

    def __init__(self, device):
        self.device = device
        self.tensor = None
        self.key = id(self)
        _arenas[self.key] = self

    def take(self):
        """the allocated buffer (an empty tensor if the library never asked); unregisters the arena"""
        _arenas.pop(self.key, None)
        if self.tensor is None:
            self.tensor = torch.empty(0, dtype=torch.uint8, device=self.device)
        return self.tensor


_arenas = {}


def _alloc(user, nbytes):
    arena = _arenas.get(user)
    try:
        arena.tensor = torch.empty(int(nbytes), dtype=torch.uint8, device=arena.device)
        return arena.tensor.data_ptr()
    except Exception:  # out of memory -> NULL, reported by the library as STP_ERR_ALLOC
        return None


_ALLOC_CB = ALLOC_FN(_alloc)  # one C thunk for every call; the arena is identified by the `user` pointer


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


STP_FORWARD_ASYNC = 4  # debug bit, include/stp_rasterizer.h
# forward passes of training steps do not wait for num_rendered (see rasterize_gaussians(async_forward=...))
ASYNC_FORWARD_DEFAULT = os.environ.get("STP_ASYNC_FORWARD", "0") == "1"
_pinned_pool = []


class NumRendered:
    """num_rendered of an ASYNCHRONOUS forward pass: the library did not wait for it (no host<->device synchronisation
    inside the forward call); it arrives in pinned host memory and is resolved on first use -- int(), ==, formatting,
    or the backward pass.  If the frame turned out not to fit the binning arena that had been sized from earlier frames
    (the image is all zeros then), resolving re-runs the forward pass synchronously into the SAME output tensors and
    replaces the three arenas (`buffers`)."""
    __slots__ = ("_value", "_pinned", "_capacity", "_rerun", "_device", "buffers", "retried")

    def __init__(self, pinned, capacity, rerun, device):
        self._value, self._pinned, self._capacity, self._rerun, self._device = None, pinned, capacity, rerun, device
        self.buffers, self.retried = None, False

    def resolve(self):
        if self._value is None:
            p = self._pinned
            # busy-wait on the pinned words (the asynchronous copy overwrites the -1 sentinels): the copy sits right
            # behind the preprocess kernel of the frame, so this returns long before the frame has finished rendering --
            # a stream / device synchronisation here would drain the whole queue and open a bubble before the backward pass
            raw = (ctypes.c_int32 * 2).from_address(p.data_ptr())
            if raw[0] < 0 or raw[1] < 0:
                import time
                t0 = time.perf_counter()
                while raw[0] < 0 or raw[1] < 0:
                    if time.perf_counter() - t0 > 30.0:
                        torch.cuda.synchronize(self._device)  # surfaces a launch failure, if that is what happened
                        if raw[0] < 0 or raw[1] < 0:
                            raise RuntimeError("asynchronous forward: num_rendered never arrived")
            R, flags = int(raw[0]), int(raw[1])
            self._pinned = None
            _pinned_pool.append(p)
            if flags & 1:  # the reference traps (auxiliary.h:226-233)
                raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
            with torch.cuda.device(self._device):
                _lib.stp_note_num_rendered(R)
            if R > self._capacity:
                out = self._rerun()
                R, self.buffers, self.retried = int(out[0]), tuple(out[3:6]), True
            self._rerun = None
            self._value = R
        return self._value

    def __int__(self):
        return self.resolve()

    __index__ = __int__

    def __eq__(self, other):
        return self.resolve() == int(other)

    def __hash__(self):
        return hash(self.resolve())

    def __repr__(self):
        return str(self.resolve())

    def __format__(self, spec):
        return format(self.resolve(), spec)

    def __del__(self):
        if self._pinned is not None and self._value is None:
            try:  # never looked at: the slot may only be reused once the copy has landed
                torch.cuda.synchronize(self._device)
                _pinned_pool.append(self._pinned)
            except Exception:
                pass


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, inv_viewprojmatrix, tan_fovx, tan_fovy, image_height, image_width, sh,
                        degree, campos, prefiltered, settings_dict, render_depth, debug, tile_band=None,
                        record_blends=True, async_forward=None, _into=None, debug_visualization=0, debug_range=None):
    """record_blends (GLOBAL / HIER modes): keep the per-pixel blend log that lets the backward pass replay the blends
    instead of sweeping the tile lists again / repeating the hierarchical re-sort; pass False for inference-only calls (GaussianRasterizer does, when no input requires a gradient).
    async_forward (default: environment STP_ASYNC_FORWARD=1): do not wait for num_rendered inside the call -- the first
    element of the result is then a NumRendered that resolves itself on first use (and re-runs the frame in the rare
    case that it outgrew the binning arena sized from earlier frames; until then the image of such a frame is black)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:68-71
    device = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    cap = BLEND_RECORD_CAP if record_blends else 0
    render_depth = bool(render_depth) or bool(debug_visualization)
    if render_depth and cap <= 0:  # the visualisations are computed from the blend log
        cap = 256
    st = settings_from_dict(settings_dict, cap, False, STP_DEBUG_DEPTH if not debug_visualization else debug_visualization,
                            debug_range) if render_depth else settings_from_dict(settings_dict, cap)
    if P == 0:  # rasterize_points.cu:93
        e8 = torch.empty(0, dtype=torch.uint8, device=device)
        return (0, torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=device),
                torch.zeros((0,), dtype=torch.int32, device=device), e8, e8.clone(), e8.clone())
    if async_forward is None:
        async_forward = ASYNC_FORWARD_DEFAULT
    async_forward = bool(async_forward) and not render_depth and not (int(debug) & 1) and _into is None
    # every pixel of the image (of the band, with tile_band) and every radius is written by the kernels
    if _into is not None:
        out_color, radii = _into
        if tile_band is not None:
            out_color.zero_()
    else:
        alloc = torch.zeros if tile_band is not None else torch.empty
        out_color = alloc((NUM_CHANNELS, H, W), dtype=torch.float32, device=device)
        radii = torch.empty((P,), dtype=torch.int32, device=device)
    geom, binning, img = _Arena(device), _Arena(device), _Arena(device)
    means3D = _f32(means3D, device)
    keep = [_f32(t, device) for t in (background, colors, opacity, scales, rotations, cov3D_precomp, viewmatrix,
                                      projmatrix, inv_viewprojmatrix, sh, campos)]
    background, colors, opacity, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, inv_viewprojmatrix, sh, campos = keep
    M = sh.size(1) if sh is not None and sh.size(0) != 0 else 0
    band = ctypes.byref(StpTileBand(int(tile_band[0]), int(tile_band[1]))) if tile_band is not None else None
    if async_forward:
        pinned = _pinned_pool.pop() if _pinned_pool else torch.empty(2, dtype=torch.int32).pin_memory()
        pinned.fill_(-1)
        n_ptr = ctypes.c_void_p(pinned.data_ptr())
    else:
        n = ctypes.c_int(0)
        n_ptr = ctypes.cast(ctypes.byref(n), ctypes.c_void_p)
    try:
        with torch.cuda.device(device):
            rc = _lib.stp_forward(_ALLOC_CB, geom.key, _ALLOC_CB, binning.key, _ALLOC_CB, img.key, P, int(degree), M,
                                  _ptr(background), W, H, ctypes.byref(st), band, _ptr(means3D), _ptr(sh), _ptr(colors),
                                  _ptr(opacity), _ptr(scales), float(scale_modifier), _ptr(rotations),
                                  _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(inv_viewprojmatrix),
                                  _ptr(campos), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
                                  out_color.data_ptr(), radii.data_ptr(),
                                  int(debug) | (STP_FORWARD_ASYNC if async_forward else 0), _stream(device), n_ptr)
    finally:
        gt, bt, it = geom.take(), binning.take(), img.take()
    if rc != 0:
        msg = _err()
        if async_forward:
            _pinned_pool.append(pinned)
        if (st.blend_record_cap > 0 and st.sort_mode != 1 and not render_depth and
                "image arena allocation failed" in msg):
            # not enough memory for the blend log: it is an optimisation (except for PPX_FULL), render without it
            torch.cuda.empty_cache()
            return rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                                       cov3D_precomp, viewmatrix, projmatrix, inv_viewprojmatrix, tan_fovx, tan_fovy,
                                       image_height, image_width, sh, degree, campos, prefiltered, settings_dict,
                                       render_depth, debug, tile_band=tile_band, record_blends=False,
                                       async_forward=async_forward, _into=_into)
        raise RuntimeError(msg)
    if not async_forward:
        return n.value, out_color, radii, gt, bt, it
    capacity = _lib.stp_binning_capacity(bt.numel(), ctypes.byref(st))

    def rerun():
        import warnings
        warnings.warn("diff_gaussian_rasterization: an asynchronous forward pass outgrew its binning arena "
                      f"(capacity {capacity}); the frame is rendered again synchronously")
        return rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                                   cov3D_precomp, viewmatrix, projmatrix, inv_viewprojmatrix, tan_fovx, tan_fovy,
                                   image_height, image_width, sh, degree, campos, prefiltered, settings_dict,
                                   render_depth, debug, tile_band=tile_band, record_blends=record_blends,
                                   async_forward=False, _into=(out_color, radii))
    # on a device without history the library falls back to the synchronous path: it then wrote R itself and marked [1]
    if int(pinned[1]) == -2:
        R = int(pinned[0])
        _pinned_pool.append(pinned)
        return R, out_color, radii, gt, bt, it
    return NumRendered(pinned, capacity, rerun, device), out_color, radii, gt, bt, it


def rasterize_gaussians_backward(background, means3D, radii, opacities, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, inv_viewprojmatrix, tan_fovx, tan_fovy,
                                 pixel_colors, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer,
                                 imageBuffer, settings_dict, debug, tile_band=None, want_param_slab=False,
                                 sync_group=None, sync_chunks=None):
    """sync_group (ours only): a torch.distributed process group over which the gradients are summed before they are
    returned (data-parallel training: views or tile bands sharded across GPUs).
    Views (no tile_band): the five PARAMETER gradients are all-reduced -- by default as ONE message, the contiguous
    parameter-gradient slab; with sync_chunks > 1 overlapped with the computation: the preprocess-backward stage runs in
    `sync_chunks` ranges of Gaussians and the all-reduce of each range's SH-gradient rows (81 % of the bytes) starts on a
    side stream as soon as the range is done (measured slower on NVSwitch, see VIEW_SYNC_CHUNKS).
    Tile bands (tile_band given): the packed screen-space accumulator (36 B/Gaussian) is all-reduced between the two
    backward stages instead (_backward_band_exchange); all eight returned gradients are then the full-frame ones."""
    if sync_chunks is None:
        sync_chunks = VIEW_SYNC_CHUNKS
    if isinstance(R, NumRendered):  # asynchronous forward: look at num_rendered now; a frame that did not fit was re-run
        R.resolve()
        if R.buffers is not None:
            geomBuffer, binningBuffer, imageBuffer = R.buffers
    device = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)  # rasterize_points.cu:169-170
    M = sh.size(1) if sh is not None and sh.numel() != 0 else 0
    st = settings_from_dict(settings_dict, blend_record_cap_of(imageBuffer, W, H))
    # ONE slab instead of nine torch::zeros (rasterize_points.cu:178-186).  Only the packed screen-space accumulator
    # (36 B/Gaussian, include/stp_rasterizer.h) is cleared; every row of the eight outputs is written by the
    # preprocess-backward kernel (zeros for culled Gaussians).  The five PARAMETER gradients come first and
    # contiguous, so a data-parallel caller can all-reduce them as one buffer (stp_sharding.py); the per-view
    # intermediates follow.
    widths = [3 * M, 3, 3, 4, 1, 3, 3, 6, 9]  # sh means3D scales rot opacity | means2D colors cov3D | accumulator
    offs = slab_offsets(P, M)  # every sub-array starts on a 16-byte boundary (128-bit stores / vector reductions)
    flat = torch.empty((offs[-1],), dtype=torch.float32, device=device)
    flat[offs[8]:offs[8] + 9 * P].zero_()
    views = [flat[offs[i]:offs[i] + widths[i] * P] for i in range(9)]
    dL_dsh, dL_dmeans3D, dL_dscales, dL_drot, dL_dopacity, dL_dmeans2D, dL_dcolors, dL_dcov3D, grad_accum = views
    param_slab = flat[:offs[5]]
    if P != 0:
        means3D = _f32(means3D, device)
        keep = [_f32(t, device) for t in (background, opacities, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                          projmatrix, inv_viewprojmatrix, pixel_colors, dL_dout_color, sh, campos)]
        (background, opacities, colors, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, inv_viewprojmatrix,
         pixel_colors, dL_dout_color, sh, campos) = keep
        radii = radii.contiguous()
        band = ctypes.byref(StpTileBand(int(tile_band[0]), int(tile_band[1]))) if tile_band is not None else None
        # R (num_rendered) is part of the reference's signature; the library re-derives the arena carve-up from the
        # size of the binning buffer instead, so a lazily resolved R never forces a host synchronisation here
        args = (P, int(degree), M, int(binningBuffer.numel()), _ptr(background), W, H, ctypes.byref(st), band,
                _ptr(means3D), _ptr(sh),
                _ptr(opacities), _ptr(colors), _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                _ptr(viewmatrix), _ptr(projmatrix), _ptr(inv_viewprojmatrix), _ptr(campos), float(tan_fovx),
                float(tan_fovy), _ptr(pixel_colors), _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer),
                _ptr(imageBuffer), _ptr(dL_dout_color), _ptr(dL_dmeans2D), _ptr(grad_accum), _ptr(dL_dopacity),
                _ptr(dL_dcolors), _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh) if M else None, _ptr(dL_dscales),
                _ptr(dL_drot), int(debug), _stream(device))
        with torch.cuda.device(device):
            if sync_group is None:
                rc = _lib.stp_backward(*args)
            elif tile_band is not None:
                rc = _backward_band_exchange(args, P, grad_accum, sync_group, device, BAND_SYNC_CHUNKS)
            elif int(sync_chunks) <= 1:
                rc = _lib.stp_backward(*args)
                if rc == 0:
                    import torch.distributed as dist
                    dist.all_reduce(param_slab, group=sync_group)  # sh | means3D | scales | rot | opacity: one message
            else:
                rc = _backward_overlapped(args, P, M, dL_dsh, flat[offs[1]:offs[5]], sync_group,
                                          int(sync_chunks), device)
        if rc != 0:
            raise RuntimeError(_err())
    grads8 = (dL_dmeans2D.view(P, 3), dL_dcolors.view(P, NUM_CHANNELS), dL_dopacity.view(P, 1), dL_dmeans3D.view(P, 3),
              dL_dcov3D.view(P, 6), dL_dsh.view(P, M, 3), dL_dscales.view(P, 3), dL_drot.view(P, 4))
    return (grads8, param_slab) if want_param_slab else grads8


def slab_offsets(P, M):
    """float offsets of the nine sub-arrays of the backward slab (sh, means3D, scales, rot, opacity | means2D, colors,
    cov3D | accumulator) and its total length; each start is rounded up to a multiple of 4 floats."""
    widths = [3 * M, 3, 3, 4, 1, 3, 3, 6, 9]
    offs, off = [], 0
    for w in widths:
        off = (off + 3) // 4 * 4
        offs.append(off)
        off += w * P
    offs.append((off + 3) // 4 * 4)
    return offs


_comm_streams = {}


def _backward_overlapped(args, P, M, dL_dsh, small, group, chunks, device):
    """render backward, then preprocess backward range by range with the all-reduce of each finished range of
    dL_dsh rows running on a side stream (NCCL enqueues behind the stream that is current when it is called)."""
    import torch.distributed as dist
    rc = _lib.stp_backward_render(*args)
    if rc != 0:
        return rc
    main = torch.cuda.current_stream(device)
    comm = _comm_streams.get(device)
    if comm is None:
        comm = _comm_streams[device] = torch.cuda.Stream(device=device)
    step = max(256, ((P + max(chunks, 1) - 1) // max(chunks, 1) + 255) // 256 * 256)
    for first in range(0, P, step):
        count = min(step, P - first)
        rc = _lib.stp_backward_preprocess(*args, first, count)
        if rc != 0:
            return rc
        ev = torch.cuda.Event()
        ev.record(main)
        if M:
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                dist.all_reduce(dL_dsh[first * 3 * M:(first + count) * 3 * M], group=group)
    with torch.cuda.stream(comm):
        comm.wait_stream(main)
        dist.all_reduce(small, group=group)
    main.wait_stream(comm)
    return 0


# view sharding: ranges of Gaussians whose SH-gradient rows are all-reduced while the next range is computed; 1 = ONE
# all-reduce of the whole contiguous parameter-gradient slab after the backward pass.  Measured on 4 B200s (C2, one view per
# rank, profiles/r02_view_exchange_chunks.txt): 1 -> 2.317 ms/step (e2e 2.74), 2 -> 2.361 (2.85), 4 -> 2.364 (2.94): the
# 0.15 ms preprocess-backward stage is too short to hide anything, and one large NVLS message moves faster than five.
VIEW_SYNC_CHUNKS = int(os.environ.get("STP_VIEW_SYNC_CHUNKS", "1"))
# ranges of Gaussians the accumulator exchange of tile-band sharding is split into (see _backward_band_exchange).
# Measured on 8 B200s (C3b, profiles/r02_band_exchange_chunks.txt): 1 range 3.22 ms/step, 2 ranges 3.44, 4 ranges 3.58 --
# one 192 MB NVLS all-reduce beats pipelining it with the 0.5 ms preprocess-backward stage in smaller messages.
BAND_SYNC_CHUNKS = int(os.environ.get("STP_BAND_SYNC_CHUNKS", "1"))


def _backward_band_exchange(args, P, grad_accum, group, device, chunks=1):
    """tile-band sharding (one view, bands of tile rows per rank): the per-Gaussian backward is linear in the packed
    screen-space gradients, and every rank holds the geometry state of every visible Gaussian (visibility does not
    depend on the band, preprocess.cu), so the ONE exchange is an all-reduce of the 36 B/Gaussian accumulator between
    the render-backward and the preprocess-backward stage -- 6.5x less than the 236 B/Gaussian of parameter gradients
     -- after which every rank finishes the same preprocess-backward and holds the full gradients.
    The accumulator can be reduced in `chunks` ranges of Gaussians on a side stream, the preprocess-backward of a range
    starting as soon as its sum has arrived; the default is ONE all-reduce (see BAND_SYNC_CHUNKS)."""
    import torch.distributed as dist
    rc = _lib.stp_backward_render(*args)
    if rc != 0:
        return rc
    main = torch.cuda.current_stream(device)
    comm = _comm_streams.get(device)
    if comm is None:
        comm = _comm_streams[device] = torch.cuda.Stream(device=device)
    step = max(256, ((P + max(chunks, 1) - 1) // max(chunks, 1) + 255) // 256 * 256)
    ranges, events = [], []
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        for first in range(0, P, step):
            count = min(step, P - first)
            if first == 0 and count == P:
                dist.all_reduce(grad_accum, group=group)  # 9 P floats, every one of them used
            else:  # the three planes of the range
                for base, width in ((0, 4), (4 * P, 4), (8 * P, 1)):
                    dist.all_reduce(grad_accum[base + first * width:base + (first + count) * width], group=group)
            ev = torch.cuda.Event()
            ev.record(comm)
            ranges.append((first, count))
            events.append(ev)
    for (first, count), ev in zip(ranges, events):
        main.wait_event(ev)
        rc = _lib.stp_backward_preprocess(*args, first, count)
        if rc != 0:
            return rc
    return 0


def blend_record_cap_of(imageBuffer, W, H):
    """blend-log capacity the forward pass allocated inside this image arena (0 = none), from its size."""
    extra = imageBuffer.numel() - _lib.stp_image_bytes(W, H, 0)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    return max(0, extra) // (tiles * 256 * 8)


def mark_visible(means3D, viewmatrix, projmatrix):
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        m, v, p = means3D.contiguous(), viewmatrix.contiguous(), projmatrix.contiguous()
        with torch.cuda.device(means3D.device):
            rc = _lib.stp_mark_visible(P, m.data_ptr(), v.data_ptr(), p.data_ptr(), present.data_ptr(),
                                       _stream(means3D.device))
        if rc != 0:
            raise RuntimeError(_err())
    return present


# ---- decoders of the opaque arenas, for the parity tests -----------------------------------------
def _wrap(ptr, owner, count, dtype):
    """torch view of `count` elements at device address `ptr` inside the uint8 tensor `owner`."""
    if not ptr:
        return None
    off = ptr - owner.data_ptr()
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    return owner[off:off + nbytes].view(dtype)


def view_geometry(geomBuffer, P, settings_dict):
    st = settings_from_dict(settings_dict)
    inv = _lib.stp_requires_cov3D_inv(ctypes.byref(st))
    v = StpGeometryView()
    _lib.stp_view_geometry(geomBuffer.data_ptr(), P, inv, ctypes.byref(v))
    f32, u8, u32 = torch.float32, torch.uint8, torch.int32
    return dict(depths=_wrap(v.depths, geomBuffer, P, f32), clamped=_wrap(v.clamped, geomBuffer, 3 * P, u8).view(P, 3),
                rects2D=_wrap(v.rects2D, geomBuffer, 2 * P, f32).view(P, 2),
                means2D=_wrap(v.means2D, geomBuffer, 2 * P, f32).view(P, 2),
                cov3D=_wrap(v.cov3D, geomBuffer, 6 * P, f32).view(P, 6),
                cov3D_inv=(_wrap(v.cov3D_inv, geomBuffer, 12 * P, f32).view(P, 3, 4) if v.cov3D_inv else None),
                conic_opacity=_wrap(v.conic_opacity, geomBuffer, 4 * P, f32).view(P, 4),
                rgb=_wrap(v.rgb, geomBuffer, 3 * P, f32).view(P, 3),
                tiles_touched=_wrap(v.tiles_touched, geomBuffer, P, u32))


def view_binning(binningBuffer, R, settings_dict=None):
    """point_list / point_list_keys (first R entries) of a binning arena.  The arena is carved for the capacity the
    forward pass allocated, which follows from its size and -- because the depth-resorting modes add the per-tile
    slabs -- from the settings; without settings both layouts are tried (only one reproduces the size exactly)."""
    cap = None
    cands = [settings_dict] if settings_dict is not None else [dict(sort_settings=dict(sort_mode=m)) for m in (0, 3)]
    for d in cands:
        st = StpSettings(int(d["sort_settings"]["sort_mode"]), *([0] * 12))
        c = _lib.stp_binning_capacity(binningBuffer.numel(), ctypes.byref(st))
        if _lib.stp_binning_bytes(c, ctypes.byref(st)) == binningBuffer.numel():
            cap = c
            break
    if cap is None:
        raise RuntimeError("binning buffer size does not match any arena layout")
    v = StpBinningView()
    _lib.stp_view_binning(binningBuffer.data_ptr(), cap, ctypes.byref(v))
    return dict(point_list=_wrap(v.point_list, binningBuffer, cap, torch.int32)[:R],
                point_list_keys=_wrap(v.point_list_keys, binningBuffer, cap, torch.int64)[:R])


def view_image(imgBuffer, W, H):
    v = StpImageView()
    _lib.stp_view_image(imgBuffer.data_ptr(), W, H, ctypes.byref(v))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    return dict(final_T=_wrap(v.final_T, imgBuffer, W * H, torch.float32).view(H, W),
                n_contrib=_wrap(v.n_contrib, imgBuffer, W * H, torch.int32).view(H, W),
                ranges=_wrap(v.ranges, imgBuffer, 2 * tiles, torch.int32).view(tiles, 2))


def last_timings():
    ms = (ctypes.c_float * 16)()
    names = (ctypes.c_char_p * 16)()
    n = _lib.stp_last_timings(ms, names, 16)
    return [(names[i].decode(), ms[i]) for i in range(n)]


def timing_summary():
    """{stage: (mean_ms, count)} over all debug&2 calls since the last summary/reset (lazy event resolve)."""
    ms = (ctypes.c_float * 16)()
    names = (ctypes.c_char_p * 16)()
    counts = (ctypes.c_int * 16)()
    n = _lib.stp_timing_summary(ms, names, counts, 16)
    return {names[i].decode(): (ms[i], counts[i]) for i in range(n)}


def timing_reset():
    _lib.stp_timing_reset()


def last_debug_stats():
    """(value at the debug pixel, min, max, mean, std) of the raw values of the last debug visualisation"""
    out = (ctypes.c_float * 5)()
    _lib.stp_last_debug_stats(out)
    return tuple(out)


def kernel_launches():
    """cumulative count of hand-written kernels launched by this thread (library kernels excluded)."""
    return int(_lib.stp_kernel_launches())
