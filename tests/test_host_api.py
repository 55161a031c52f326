"""Host-side mirror of the reference's Python package (diff_gaussian_rasterization/__init__.py):
names, field order, defaults, dict schema and error behaviour -- no GPU needed."""
import json

import pytest
import torch

import diff_gaussian_rasterization as dgr
from diff_gaussian_rasterization import (CullingSettings, ExtendedSettings, GaussianRasterizationSettings,
                                         GaussianRasterizer, GlobalSortOrder, SortMode, SortQueueSizes, SortSettings)


def test_enums_match_reference_values():  # __init__.py:175-191, rasterizer.h:27-41
    assert [m.value for m in SortMode] == [0, 1, 2, 3]
    assert [m.name for m in SortMode] == ["GLOBAL", "PPX_FULL", "PPX_KBUFFER", "HIER"]
    assert [m.name for m in GlobalSortOrder] == ["Z_DEPTH", "DISTANCE", "PTD_CENTER", "PTD_MAX"]
    assert str(SortMode.HIER) == "HIER"


def test_defaults_and_dict_schema():  # __init__.py:193-233, rasterizer.h:137-158
    d = ExtendedSettings().to_dict()
    assert d == {
        "sort_settings": {"queue_sizes": {"tile_4x4": 64, "tile_2x2": 8, "per_pixel": 4}, "sort_mode": 0, "sort_order": 0},
        "culling_settings": {"rect_bounding": False, "tight_opacity_bounding": False, "tile_based_culling": False,
                             "hierarchical_4x4_culling": False},
        "load_balancing": False, "proper_ewa_scaling": False}
    assert json.loads(ExtendedSettings().to_json()) == d
    # enums are serialised as ints
    s = ExtendedSettings(sort_settings=SortSettings(sort_mode=SortMode.HIER, sort_order=GlobalSortOrder.PTD_MAX))
    assert s.to_dict()["sort_settings"]["sort_mode"] == 3 and s.to_dict()["sort_settings"]["sort_order"] == 3


def test_instances_do_not_share_mutable_defaults():
    a, b = ExtendedSettings(), ExtendedSettings()
    a.sort_settings.queue_sizes.per_pixel = 16
    assert b.sort_settings.queue_sizes.per_pixel == 4


def test_from_dict_roundtrip_and_set_value(tmp_path):  # __init__.py:236-246
    s = ExtendedSettings.from_dict({"sort_settings": {"sort_mode": 3, "queue_sizes": {"per_pixel": 8}},
                                    "culling_settings": {"tile_based_culling": True}})
    assert s.sort_settings.sort_mode is SortMode.HIER and s.sort_settings.queue_sizes.per_pixel == 8
    assert s.sort_settings.queue_sizes.tile_2x2 == 8 and s.culling_settings.tile_based_culling is True
    assert ExtendedSettings.from_dict(s.to_dict()) == s
    p = tmp_path / "preset.json"
    p.write_text(s.to_json())
    assert ExtendedSettings.from_json(str(p)) == s
    s.set_value("load_balancing", True)
    s.set_value("rect_bounding", True)
    s.set_value("sort_order", GlobalSortOrder.PTD_CENTER)
    s.set_value("tile_2x2", 20)
    assert s.load_balancing and s.culling_settings.rect_bounding
    assert s.sort_settings.sort_order == GlobalSortOrder.PTD_CENTER and s.sort_settings.queue_sizes.tile_2x2 == 20


def test_raster_settings_field_order():  # __init__.py:248-263
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "inv_viewprojmatrix", "sh_degree", "campos", "prefiltered", "settings", "render_depth", "debug")


def _rs():
    e = torch.eye(4)
    return GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, e, e, e, 0, torch.zeros(3), False,
                                         ExtendedSettings(), False, False)


def test_forward_argument_validation():  # __init__.py:285-289
    r = GaussianRasterizer(_rs())
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), colors_precomp=m)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_native_module_surface():  # ext.cpp:15-19
    for n in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert callable(getattr(dgr._C, n))
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):  # rasterize_points.cu:68-71
        dgr._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4), *([torch.zeros(0)] * 4), 1.0, torch.zeros(0),
                                   *([torch.eye(4)] * 3), 1.0, 1.0, 16, 16, torch.zeros(0), 0, torch.zeros(3), False,
                                   ExtendedSettings().to_dict(), False, False)


def test_settings_dict_keys_are_mandatory():  # rasterizer.h:160-182 uses .at()
    d = ExtendedSettings().to_dict()
    del d["load_balancing"]
    with pytest.raises(KeyError):
        dgr._C.settings_from_dict(d)


def test_backward_slab_layout_is_aligned_and_matches_sharding_helper():
    """every sub-array of the backward slab starts on a 16-byte boundary for any P (vector reductions / float4 stores),
    the five parameter gradients form one contiguous prefix, and stp_sharding.split_param_slab reads the same layout"""
    import stp_sharding as sh
    from diff_gaussian_rasterization import _C
    for P, M in ((1, 16), (3, 16), (7, 4), (1025, 1), (6001, 9), (4096, 16), (5, 0)):
        offs = _C.slab_offsets(P, M)
        widths = [3 * M, 3, 3, 4, 1, 3, 3, 6, 9]
        assert all(o % 4 == 0 for o in offs)
        assert all(offs[i] + widths[i] * P <= offs[i + 1] for i in range(9))
        assert offs[5] == sh.param_slab_numel(P, M)
        slab = torch.arange(offs[5], dtype=torch.float32)
        m3, shg, op, sc, ro = sh.split_param_slab(slab, P, M)
        assert shg.shape == (P, M, 3) and m3.shape == (P, 3) and op.shape == (P, 1)
        if M:
            assert shg.reshape(-1)[0].item() == offs[0]
        assert m3.reshape(-1)[0].item() == offs[1] and sc.reshape(-1)[0].item() == offs[2]
        assert ro.reshape(-1)[0].item() == offs[3] and op.reshape(-1)[0].item() == offs[4]


def test_blend_log_capacity_is_recovered_from_the_arena_size():
    """backward learns the capacity of the forward pass's blend log from the size of the image arena alone"""
    from diff_gaussian_rasterization import _C
    for W, H in ((64, 48), (1920, 1080), (17, 5)):
        for cap in (0, 1, 6, 256):
            nbytes = _C._lib.stp_image_bytes(W, H, cap)

            class _Buf:  # only numel() is consulted; no need to allocate gigabytes here
                def numel(self):
                    return nbytes
            assert _C.blend_record_cap_of(_Buf(), W, H) == cap
