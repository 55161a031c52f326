"""Drop-in ``diff_gaussian_rasterization`` for the StopThePop training loop, backed by the B200-native
CUDA library (libstp_rasterizer.so, sm_100a).

Public surface = the reference's Python package (diff_gaussian_rasterization/__init__.py of
r4dl/StopThePop-Rasterization): same names, field order, defaults, argument meaning and exceptions.

    GaussianRasterizationSettings  (NamedTuple, __init__.py:248-263)
    GaussianRasterizer             (nn.Module,  __init__.py:265-314)
    rasterize_gaussians / _RasterizeGaussians (autograd.Function, __init__.py:32-172)
    SortMode, GlobalSortOrder      (IntEnum,    __init__.py:175-191)
    SortQueueSizes, SortSettings, CullingSettings, ExtendedSettings (dataclasses, __init__.py:193-246)

Differences that are deliberate and invisible to callers: dataclass defaults use default_factory
(the reference's mutable defaults do not import on Python >= 3.11), ``ExtendedSettings.from_dict``
does not need ``dacite``, and the native module is a ctypes binding of a C ABI instead of pybind11.
"""
import json
from dataclasses import asdict, dataclass, field
from enum import IntEnum
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C


def enum_dict_factory(data):
    def convert_value(obj):
        if isinstance(obj, IntEnum):
            return obj.value
        return obj
    return dict((k, convert_value(v)) for k, v in data)


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(
    means3D,
    means2D,
    sh,
    colors_precomp,
    opacities,
    scales,
    rotations,
    cov3Ds_precomp,
    raster_settings,
    tile_band=None,
    sync_group=None,
    async_forward=None,
):
    # The blend log (2 KB per pixel) only pays off when a backward pass will follow.  Decided HERE, not inside
    # Function.forward: there grad mode is always off and ctx.needs_input_grad reflects requires_grad even under
    # torch.no_grad(), so evaluation renders of a trained model would record (and allocate) the log for nothing.
    record_blends = torch.is_grad_enabled() and any(
        isinstance(t, torch.Tensor) and t.requires_grad
        for t in (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
    return _RasterizeGaussians.apply(
        means3D,
        means2D,
        sh,
        colors_precomp,
        opacities,
        scales,
        rotations,
        cov3Ds_precomp,
        raster_settings,
        tile_band,
        sync_group,
        record_blends,
        async_forward,
    )


class _RasterizeGaussians(torch.autograd.Function):
    """autograd glue; argument packing follows __init__.py:70-93 (forward) and :122-158 (backward)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, tile_band=None, sync_group=None, record_blends=None, async_forward=None):
        # tile_band (ours only, multi-GPU tile sharding): (row0, row1) of 16-pixel tile rows this rank renders
        # sync_group (ours only): process group over which backward sums the parameter gradients (overlapped exchange)
        # async_forward (ours only): do not wait for num_rendered inside the forward call (_C.NumRendered)
        rs = raster_settings
        if record_blends is None:  # direct .apply callers: the wrapper above knows better (grad mode)
            record_blends = any(ctx.needs_input_grad)
        args = (
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.inv_viewprojmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
            rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.settings.to_dict(), rs.render_depth,
            rs.debug,
        )
        ctx.settings_dict = args[19]  # backward uses the settings the forward pass ran with
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
            try:
                num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(
                    *args, record_blends=record_blends, tile_band=tile_band, async_forward=async_forward)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(
                *args, record_blends=record_blends, tile_band=tile_band, async_forward=async_forward)

        ctx.raster_settings = rs
        ctx.tile_band = tile_band
        ctx.sync_group = sync_group
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, opacities, scales, rotations, cov3Ds_precomp, radii, sh, color,
                              geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        num_rendered = ctx.num_rendered
        rs = ctx.raster_settings
        (colors_precomp, means3D, opacities, scales, rotations, cov3Ds_precomp, radii, sh, color, geomBuffer,
         binningBuffer, imgBuffer) = ctx.saved_tensors
        # (an asynchronous forward pass hands over a lazily resolved num_rendered; the native call below resolves it and,
        # should the frame have outgrown its binning arena, uses the buffers of the re-run instead of the saved ones)
        args = (rs.bg, means3D, radii, opacities, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.inv_viewprojmatrix, rs.tanfovx, rs.tanfovy, color, grad_out_color, sh,
                rs.sh_degree, rs.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer, ctx.settings_dict,
                rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                grads8 = _C.rasterize_gaussians_backward(*args, tile_band=ctx.tile_band, sync_group=ctx.sync_group)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads8 = _C.rasterize_gaussians_backward(*args, tile_band=ctx.tile_band, sync_group=ctx.sync_group)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = grads8
        return (
            grad_means3D,
            grad_means2D,
            grad_sh,
            grad_colors_precomp,
            grad_opacities,
            grad_scales,
            grad_rotations,
            grad_cov3Ds_precomp,
            None,
            None,
            None,
            None,
            None,
        )


class SortMode(IntEnum):
    GLOBAL = 0
    PPX_FULL = 1
    PPX_KBUFFER = 2
    HIER = 3

    def __str__(self):
        return self.name


class GlobalSortOrder(IntEnum):
    Z_DEPTH = 0
    DISTANCE = 1
    PTD_CENTER = 2
    PTD_MAX = 3

    def __str__(self):
        return self.name


@dataclass
class SortQueueSizes:
    tile_4x4: int = 64
    tile_2x2: int = 8
    per_pixel: int = 4

    def set_value(self, key, value):
        if key in self.__dataclass_fields__.keys():
            self.__setattr__(key, value)


@dataclass
class SortSettings:
    queue_sizes: SortQueueSizes = field(default_factory=SortQueueSizes)
    sort_mode: SortMode = SortMode.GLOBAL
    sort_order: GlobalSortOrder = GlobalSortOrder.Z_DEPTH

    def set_value(self, key, value):
        if key in self.__dataclass_fields__.keys():
            self.__setattr__(key, value)
        else:
            self.queue_sizes.set_value(key, value)


@dataclass
class CullingSettings:
    rect_bounding: bool = False
    tight_opacity_bounding: bool = False
    tile_based_culling: bool = False
    hierarchical_4x4_culling: bool = False

    def set_value(self, key, value):
        if key in self.__dataclass_fields__.keys():
            self.__setattr__(key, value)


@dataclass
class ExtendedSettings:
    sort_settings: SortSettings = field(default_factory=SortSettings)
    culling_settings: CullingSettings = field(default_factory=CullingSettings)
    load_balancing: bool = False
    proper_ewa_scaling: bool = False

    def to_dict(self):
        # same result as asdict(self, dict_factory=enum_dict_factory) (the reference, __init__.py:231), written out:
        # this runs twice per training step and the generic recursive asdict costs ~25 us
        ss, cs, q = self.sort_settings, self.culling_settings, self.sort_settings.queue_sizes
        val = lambda v: v.value if isinstance(v, IntEnum) else v  # noqa: E731
        return {
            "sort_settings": {
                "queue_sizes": {"tile_4x4": val(q.tile_4x4), "tile_2x2": val(q.tile_2x2), "per_pixel": val(q.per_pixel)},
                "sort_mode": val(ss.sort_mode), "sort_order": val(ss.sort_order)},
            "culling_settings": {
                "rect_bounding": val(cs.rect_bounding), "tight_opacity_bounding": val(cs.tight_opacity_bounding),
                "tile_based_culling": val(cs.tile_based_culling),
                "hierarchical_4x4_culling": val(cs.hierarchical_4x4_culling)},
            "load_balancing": val(self.load_balancing),
            "proper_ewa_scaling": val(self.proper_ewa_scaling),
        }

    def to_json(self):
        return json.dumps(self.to_dict())

    @staticmethod
    def from_dict(dict):
        # what dacite.from_dict(..., Config(cast=[IntEnum])) does in the reference (__init__.py:236-237):
        # missing keys fall back to the dataclass defaults, ints are cast to the enums.
        d = dict
        ss = d.get("sort_settings", {})
        q = ss.get("queue_sizes", {})
        cs = d.get("culling_settings", {})
        qd, sd, cd, ed = SortQueueSizes(), SortSettings(), CullingSettings(), ExtendedSettings()
        return ExtendedSettings(
            sort_settings=SortSettings(
                queue_sizes=SortQueueSizes(tile_4x4=int(q.get("tile_4x4", qd.tile_4x4)),
                                           tile_2x2=int(q.get("tile_2x2", qd.tile_2x2)),
                                           per_pixel=int(q.get("per_pixel", qd.per_pixel))),
                sort_mode=SortMode(ss.get("sort_mode", sd.sort_mode)),
                sort_order=GlobalSortOrder(ss.get("sort_order", sd.sort_order))),
            culling_settings=CullingSettings(
                rect_bounding=bool(cs.get("rect_bounding", cd.rect_bounding)),
                tight_opacity_bounding=bool(cs.get("tight_opacity_bounding", cd.tight_opacity_bounding)),
                tile_based_culling=bool(cs.get("tile_based_culling", cd.tile_based_culling)),
                hierarchical_4x4_culling=bool(cs.get("hierarchical_4x4_culling", cd.hierarchical_4x4_culling))),
            load_balancing=bool(d.get("load_balancing", ed.load_balancing)),
            proper_ewa_scaling=bool(d.get("proper_ewa_scaling", ed.proper_ewa_scaling)))

    @staticmethod
    def from_json(json_filename):
        return ExtendedSettings.from_dict(json.load(open(json_filename)))

    def set_value(self, key, value):
        if key in self.__dataclass_fields__.keys():
            self.__setattr__(key, value)
        else:
            self.culling_settings.set_value(key, value)
            self.sort_settings.set_value(key, value)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    inv_viewprojmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    settings: ExtendedSettings
    render_depth: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, tile_band=None, sync_group=None, async_forward=None):
        """tile_band (ours only): (row0, row1) band of 16-pixel tile rows rendered / differentiated by this rank when
        one view is sharded across GPUs (stp_sharding.py); None = the whole image, as in the reference.
        sync_group (ours only): torch.distributed process group; backward returns parameter gradients already summed
        over the group, with the all-reduce overlapped with the preprocess-backward stage (_C.py)."""
        super().__init__()
        self.raster_settings = raster_settings
        self.tile_band = tile_band
        self.sync_group = sync_group
        self.async_forward = async_forward  # None: environment STP_ASYNC_FORWARD (default off); see _C.NumRendered

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        # Invoke the CUDA rasterization routine
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings, self.tile_band, self.sync_group, self.async_forward)
