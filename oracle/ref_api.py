"""Oracle harness: the reference's own CUDA build (oracle/_ref/_C*.so) behind a small Python API.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py (--impl reference /
cpu_baseline) may import this module; the product (stopthepop-rasterization_b200/) never does.

The reference package's __init__.py cannot be imported on this image (needs `dacite`, and its
mutable dataclass defaults are rejected by Python >= 3.11), so the compiled pybind module
is loaded directly and called with a hand-built settings dict (schema rasterizer.h:137-158).
Decoders for the reference's opaque buffers follow rasterizer_impl.cu:175-217 (each sub-array
aligned to 128 B of the absolute address).
"""
import glob
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def available():
    return bool(glob.glob(os.path.join(_HERE, "_ref", "_C*.so")))


def module():
    global _mod
    if _mod is None:
        cands = glob.glob(os.path.join(_HERE, "_ref", "_C*.so"))
        if not cands:
            raise ImportError("oracle/_ref/_C*.so missing: run python oracle/build_ref.py where /root/reference exists")
        spec = importlib.util.spec_from_file_location("_C", cands[0])
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def _e(device):
    return torch.empty(0, dtype=torch.float32, device=device)


def forward(scene, cam, settings_dict, colors_precomp=None, cov3D_precomp=None, scale_modifier=1.0, debug=False,
            render_depth=False):
    """scene/cam: stp_scenes.Scene/Camera already on the GPU.  Returns the reference's 6-tuple."""
    dev = scene.means3D.device
    C = module()
    use_cov = cov3D_precomp is not None
    return C.rasterize_gaussians(
        cam.bg, scene.means3D, colors_precomp if colors_precomp is not None else _e(dev), scene.opacities,
        _e(dev) if use_cov else scene.scales, _e(dev) if use_cov else scene.rotations, float(scale_modifier),
        cov3D_precomp if use_cov else _e(dev), cam.viewmatrix, cam.projmatrix, cam.inv_viewprojmatrix,
        cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width,
        _e(dev) if colors_precomp is not None else scene.shs, scene.sh_degree, cam.campos, False, settings_dict,
        bool(render_depth), debug)


def backward(scene, cam, settings_dict, fwd_out, dL_dout, colors_precomp=None, cov3D_precomp=None, scale_modifier=1.0,
             debug=False):
    dev = scene.means3D.device
    C = module()
    R, color, radii, geom, binning, img = fwd_out
    use_cov = cov3D_precomp is not None
    return C.rasterize_gaussians_backward(
        cam.bg, scene.means3D, radii, scene.opacities, colors_precomp if colors_precomp is not None else _e(dev),
        _e(dev) if use_cov else scene.scales, _e(dev) if use_cov else scene.rotations, float(scale_modifier),
        cov3D_precomp if use_cov else _e(dev), cam.viewmatrix, cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx,
        cam.tanfovy, color, dL_dout, _e(dev) if colors_precomp is not None else scene.shs, scene.sh_degree, cam.campos,
        geom, R, binning, img, settings_dict, debug)


def requires_inv(settings_dict):
    ss = settings_dict["sort_settings"]
    return ss["sort_mode"] != 0 or ss["sort_order"] in (2, 3)


class _Carver:
    def __init__(self, buf):
        self.buf = buf
        self.base = buf.data_ptr()
        self.cur = self.base

    def take(self, count, dtype):
        es = torch.empty((), dtype=dtype).element_size()
        start = (self.cur + 127) & ~127
        off = start - self.base
        self.cur = start + count * es
        return self.buf[off:off + count * es].view(dtype)


def decode_geometry(geom, P, settings_dict):
    """fields that precede the CUB scan temp (whose size is CUB-version dependent)."""
    c = _Carver(geom)
    out = {}
    out["depths"] = c.take(P, torch.float32)
    out["clamped"] = c.take(3 * P, torch.uint8).view(P, 3)
    out["internal_radii"] = c.take(P, torch.int32)
    out["rects2D"] = c.take(2 * P, torch.float32).view(P, 2)
    out["means2D"] = c.take(2 * P, torch.float32).view(P, 2)
    out["cov3D"] = c.take(6 * P, torch.float32).view(P, 6)
    if requires_inv(settings_dict):
        out["cov3D_inv"] = c.take(12 * P, torch.float32).view(P, 3, 4)
    out["conic_opacity"] = c.take(4 * P, torch.float32).view(P, 4)
    out["rgb"] = c.take(3 * P, torch.float32).view(P, 3)
    out["tiles_touched"] = c.take(P, torch.int32)
    return out


def decode_binning(binning, R):
    c = _Carver(binning)
    out = {}
    out["point_list"] = c.take(R, torch.int32)
    out["point_list_unsorted"] = c.take(R, torch.int32)
    out["point_list_keys"] = c.take(R, torch.int64)
    out["point_list_keys_unsorted"] = c.take(R, torch.int64)
    return out


def decode_image(img, W, H):
    c = _Carver(img)
    N = W * H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out = {}
    out["final_T"] = c.take(N, torch.float32).view(H, W)
    out["n_contrib"] = c.take(N, torch.int32).view(H, W)
    out["ranges"] = c.take(2 * N, torch.int32)[:2 * tiles].view(tiles, 2)
    return out
