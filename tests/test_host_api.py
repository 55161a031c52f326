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
