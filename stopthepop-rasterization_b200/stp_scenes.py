"""Seeded synthetic Gaussian clouds and cameras for parity tests, smoke() and bench.py.

Definition: SURVEY.md section 8(d) ("Synthetic inputs").  Everything is generated on the CPU in
float32 from torch.Generator().manual_seed(1000 + config_id) and moved to the device afterwards, so
the reference build and this library always see bit-identical inputs.  The upstream gradient is
N(0,1)[3,H,W] from seed 2000 + config_id.
"""
import math
from typing import NamedTuple

import torch


class Scene(NamedTuple):
    means3D: torch.Tensor      # [P,3]
    scales: torch.Tensor       # [P,3]
    rotations: torch.Tensor    # [P,4]  normalised (r,x,y,z)
    opacities: torch.Tensor    # [P,1]
    shs: torch.Tensor          # [P,16,3]
    sh_degree: int


class Camera(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor          # [4,4] = W2C^T
    projmatrix: torch.Tensor          # [4,4] = (P W2C)^T
    inv_viewprojmatrix: torch.Tensor  # [4,4] = inverse(projmatrix)
    campos: torch.Tensor              # [3]
    bg: torch.Tensor                  # [3]


# name -> (config_id, P, W, H)
CONFIGS = {
    "C1": (1, 1_000, 256, 256),
    "C2": (2, 1_000_000, 1920, 1080),
    "C3": (3, 4_000_000, 1920, 1080),
    "C4": (4, 4_000_000, 3840, 2160),
    "C5": (5, 10_000_000, 1920, 1080),
}


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float64)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float64)


def make_camera(W, H, yaw=0.0):
    """pin-hole camera of SURVEY 8(d); `yaw` (rad) rotates the rig about world-Y (config 5: k*0.08)."""
    focal = 0.9 * W
    tanfovx = 0.5 * W / focal
    tanfovy = 0.5 * H / focal
    R = _rot_x(0.05) @ _rot_y(0.10) @ _rot_y(yaw)
    t = torch.tensor([0.1, -0.2, 0.3], dtype=torch.float64)
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = R
    w2c[:3, 3] = t
    znear, zfar = 0.01, 100.0
    top, right = tanfovy * znear, tanfovx * znear
    Pm = torch.zeros(4, 4, dtype=torch.float64)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    view = w2c.t().contiguous().float()
    proj = (Pm @ w2c).t().contiguous().float()
    inv = torch.linalg.inv(proj.double()).float().contiguous()
    campos = (-(R.t() @ t)).float()
    bg = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float32)
    return Camera(H, W, float(tanfovx), float(tanfovy), view, proj, inv, campos, bg), R, t


def make_scene(P, W, H, seed, sigma_scale=None):
    """P Gaussians distributed through (and a little outside) the frustum of make_camera(W, H).
    sigma_scale overrides the pixel-footprint factor H/1080 (used by the small golden fixtures to get
    long per-tile lists on tiny images)."""
    g = torch.Generator().manual_seed(seed)
    cam, R, t = make_camera(W, H)
    focal = 0.9 * W

    def U(lo, hi, *shape):
        return lo + (hi - lo) * torch.rand(*shape, generator=g, dtype=torch.float32)

    z = U(-1.0, 30.0, P)
    x = z * cam.tanfovx * U(-1.15, 1.15, P)
    y = z * cam.tanfovy * U(-1.15, 1.15, P)
    p_cam = torch.stack([x, y, z], dim=1).double()
    means3D = ((p_cam - t) @ R).float().contiguous()  # R^T (p - t), row-vector form
    sigma_px = torch.exp(U(math.log(0.4), math.log(6.0), P)) * (H / 1080.0 if sigma_scale is None else sigma_scale)
    scales = (sigma_px * z.abs() / focal).unsqueeze(1) * torch.exp(U(math.log(0.3), 0.0, P, 3))
    scales = scales.clamp_min(1e-7).contiguous()
    q = torch.randn(P, 4, generator=g, dtype=torch.float32)
    rotations = (q / q.norm(dim=1, keepdim=True)).contiguous()
    opacities = torch.sigmoid(2.0 * torch.randn(P, 1, generator=g, dtype=torch.float32)).contiguous()
    shs = 0.1 * torch.randn(P, 16, 3, generator=g, dtype=torch.float32)
    shs[:, 0, :] = U(-1.0, 1.5, P, 3)
    return Scene(means3D, scales, rotations, opacities, shs.contiguous(), 3), cam


def make_config(name, P=None, W=None, H=None):
    cid, p0, w0, h0 = CONFIGS[name]
    return make_scene(P or p0, W or w0, H or h0, 1000 + cid)


def make_upstream_grad(W, H, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(3, H, W, generator=g, dtype=torch.float32)


def to_device(obj, device):
    return type(obj)(*[v.to(device) if isinstance(v, torch.Tensor) else v for v in obj])


def default_settings_dict(**kw):
    """settings dict with the schema of rasterizer.h:137-158; keyword overrides are flat like
    ExtendedSettings.set_value."""
    d = {
        "sort_settings": {"sort_mode": 0, "sort_order": 0,
                          "queue_sizes": {"tile_4x4": 64, "tile_2x2": 8, "per_pixel": 4}},
        "culling_settings": {"rect_bounding": False, "tight_opacity_bounding": False, "tile_based_culling": False,
                             "hierarchical_4x4_culling": False},
        "load_balancing": False,
        "proper_ewa_scaling": False,
    }
    for k, v in kw.items():
        if k in d:
            d[k] = v
        elif k in d["sort_settings"]:
            d["sort_settings"][k] = int(v)
        elif k in d["sort_settings"]["queue_sizes"]:
            d["sort_settings"]["queue_sizes"][k] = int(v)
        elif k in d["culling_settings"]:
            d["culling_settings"][k] = bool(v)
        else:
            raise KeyError(k)
    return d


STOPTHEPOP_PRESET = dict(sort_mode=3, sort_order=3, rect_bounding=True, tight_opacity_bounding=True,
                         tile_based_culling=True, hierarchical_4x4_culling=True, load_balancing=True,
                         proper_ewa_scaling=False)
