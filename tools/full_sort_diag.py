"""diagnostic: PPX_FULL at full size, ours vs the reference build: which pixels differ and why (run on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stopthepop-rasterization_b200")):
    sys.path.insert(0, p)
import torch
import stp_scenes as S
from diff_gaussian_rasterization import _C
from oracle import ref_api as ref
dev = torch.device("cuda:0")
scene = sys.argv[1] if len(sys.argv) > 1 else "C4"
sc, cam = S.make_config(scene)
sc, cam = S.to_device(sc, dev), S.to_device(cam, dev)
W, H = cam.image_width, cam.image_height
d = S.default_settings_dict(sort_mode=1)
e = torch.empty(0, device=dev)
out = _C.rasterize_gaussians(cam.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e, cam.viewmatrix,
                             cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx, cam.tanfovy, H, W, sc.shs, 3, cam.campos,
                             False, d, False, False)
rr = ref.forward(sc, cam, d)
rr2 = ref.forward(sc, cam, d)
print("reference deterministic:", torch.equal(rr[1], rr2[1]))
diff = (rr[1] - out[1]).abs().amax(0)
scale = rr[1].abs().max().item()
bad = torch.nonzero(diff > 1e-5 * scale)
img_o, img_r = _C.view_image(out[5], W, H), ref.decode_image(rr[5], W, H)
ranges = img_o["ranges"]
gx = (W + 15) // 16
print("differing pixels:", bad.shape[0])
for y, x in bad.tolist()[:20]:
    t = (y // 16) * gx + x // 16
    n = int(ranges[t, 1] - ranges[t, 0])
    print(f"pixel ({x},{y}) tile {t} len {n} diff {diff[y, x].item():.3e} n_contrib ours {int(img_o['n_contrib'][y, x])} ref {int(img_r['n_contrib'][y, x])} "
          f"T ours {img_o['final_T'][y, x].item():.6f} ref {img_r['final_T'][y, x].item():.6f} ours {out[1][:, y, x].tolist()} ref {rr[1][:, y, x].tolist()}")
nc = (img_o["n_contrib"] != img_r["n_contrib"]).sum().item()
print("n_contrib mismatches:", nc)
# near-ties: depth along the pixel's ray (float64) of the entries of the pixel's tile, smallest gap between neighbours
g = _C.view_geometry(out[3], sc.means3D.shape[0], d)
pl = _C.view_binning(out[4], out[0], d)["point_list"].long()
inv = g["cov3D_inv"].double()
ivp = cam.inv_viewprojmatrix.double()
cp = cam.campos.double()
for y, x in bad.tolist()[:6]:
    t = (y // 16) * gx + x // 16
    ids = pl[int(ranges[t, 0]):int(ranges[t, 1])]
    # pix2world (auxiliary.h:71-81): ndc = (2 p + 1)/S - 1 ... done in float64 here; only the near-tie structure matters
    ndc = torch.tensor([2.0 * x / W - 1.0 + 1.0 / W * 0, 2.0 * y / H - 1.0, 1.0, 1.0], dtype=torch.float64, device=dev)
    ndc[0] = (2.0 * x + 1.0) / W - 1.0
    ndc[1] = (2.0 * y + 1.0) / H - 1.0
    wpos = ndc @ ivp
    wpos = wpos[:3] / wpos[3]
    ray = wpos - cp
    ray = ray / ray.norm()
    S6 = inv[ids]
    a, b, c = S6[:, 0, :3], S6[:, 1, :3], S6[:, 2, :3]
    vx = a[:, 0] * ray[0] + a[:, 1] * ray[1] + a[:, 2] * ray[2]
    vy = a[:, 1] * ray[0] + b[:, 0] * ray[1] + b[:, 1] * ray[2]
    vz = a[:, 2] * ray[0] + b[:, 1] * ray[1] + b[:, 2] * ray[2]
    num = c[:, 0] * ray[0] + c[:, 1] * ray[1] + c[:, 2] * ray[2]
    den = (vx * ray[0] + vy * ray[1] + vz * ray[2]).clamp_min(1e-5)
    dep = num / den
    srt, _ = torch.sort(dep)
    gaps = ((srt[1:] - srt[:-1]) / srt[1:].abs().clamp_min(1e-9)).abs()
    k = torch.argsort(gaps)[:3]
    print(f"pixel ({x},{y}): smallest relative depth gaps {gaps[k].tolist()} (float32 ulp ~6e-8)")
