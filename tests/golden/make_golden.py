#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference build
(oracle/_ref/_C*.so) on a GPU.  The reference ships no tests or golden vectors,
so these outputs of the reference itself are the pins of the oracle and of the CUDA path.

    gpurun -- python tests/golden/make_golden.py --out gpurun_out/golden      # on the B200 box
    cp gpurun_out/golden/*.npz tests/golden/                                   # here, then commit

Each fixture stores the inputs (so no RNG reproducibility is assumed), the settings dict and every
output the reference exposes: R, radii, out_color, the decoded geometry / binning / image buffers and,
where the reference implements it, the eight gradients plus a second run of them (noise floor).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stopthepop-rasterization_b200"))
sys.path.insert(0, ROOT)
import stp_scenes as S  # noqa: E402
from oracle import ref_api as ref  # noqa: E402

D = S.default_settings_dict
# scenes: name -> (P, W, H, seed, sigma_scale).  Inputs are stored once per scene.
SCENES = {
    "A": (1500, 100, 72, 11, 2.0),       # ~250-entry tile lists, ragged image edges (100x72 = 7x5 tiles)
    "B": (300, 100, 72, 20, 1.5),        # sparse: lists shorter than one 32-entry batch
    "L": (9000, 64, 48, 24, 2.0),        # long lists (> 1024 per tile) for the windowed full sort
    "C1": (1000, 256, 256, 1001, None),  # BASELINE.json configs[0]
}
CASES = [
    # name, scene, sh_degree, settings, backward?
    ("global_default", "A", 3, D(), True),
    ("global_distance_ewa", "A", 1, D(sort_order=1, proper_ewa_scaling=True), True),
    ("global_tbc_ptdmax", "A", 1,
     D(sort_order=3, rect_bounding=True, tight_opacity_bounding=True, tile_based_culling=True), True),
    ("global_ptdcenter_lb", "A", 1, D(sort_order=2, load_balancing=True), True),
    ("hier_default", "A", 1, D(sort_mode=3), True),
    ("hier_preset", "A", 1, D(**S.STOPTHEPOP_PRESET), True),
    ("hier_cull_only", "A", 1, D(sort_mode=3, hierarchical_4x4_culling=True), True),
    ("hier_q16_20", "A", 1, D(sort_mode=3, per_pixel=16, tile_2x2=20), True),
    ("hier_q8_12", "A", 1, D(sort_mode=3, per_pixel=8, tile_2x2=12), True),
    ("hier_sparse", "B", 1, D(sort_mode=3), True),
    ("hier_long", "L", 0, D(sort_mode=3, hierarchical_4x4_culling=True), False),
    ("kbuffer16", "A", 1, D(sort_mode=2, per_pixel=16), True),
    ("kbuffer4", "A", 1, D(sort_mode=2, per_pixel=4), True),
    ("full_sort", "A", 1, D(sort_mode=1), False),
    ("full_sort_long", "L", 0, D(sort_mode=1), False),
    ("c1_config", "C1", 3, D(), False),
]


def npy(t):
    return t.detach().cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    dev = torch.device("cuda:0")
    scenes = {}
    for sname, (P, W, H, seed, sig) in SCENES.items():
        sc_c, cam_c = S.make_scene(P, W, H, seed, sigma_scale=sig)
        scenes[sname] = (sc_c, cam_c, seed)
        np.savez_compressed(
            os.path.join(a.out, f"scene_{sname}.npz"), P=P, W=W, H=H, seed=seed,
            means3D=npy(sc_c.means3D), scales=npy(sc_c.scales), rotations=npy(sc_c.rotations),
            opacities=npy(sc_c.opacities), shs=npy(sc_c.shs), viewmatrix=npy(cam_c.viewmatrix),
            projmatrix=npy(cam_c.projmatrix), inv_viewprojmatrix=npy(cam_c.inv_viewprojmatrix),
            campos=npy(cam_c.campos), bg=npy(cam_c.bg), tanfovx=cam_c.tanfovx, tanfovy=cam_c.tanfovy,
            dL_dout=npy(S.make_upstream_grad(W, H, seed + 1000)))
    for name, sname, deg, settings, bwd in CASES:
        sc_c, cam_c, seed = scenes[sname]
        M = (deg + 1) ** 2
        sc_c = sc_c._replace(shs=sc_c.shs[:, :M, :].contiguous(), sh_degree=deg)
        P, W, H = sc_c.means3D.shape[0], cam_c.image_width, cam_c.image_height
        sc, cam = S.to_device(sc_c, dev), S.to_device(cam_c, dev)
        out = ref.forward(sc, cam, settings)
        torch.cuda.synchronize()
        R, color, radii, geom, binning, img = out
        fx = dict(settings=json.dumps(settings), scene=sname, sh_degree=deg, R=R, out_color=npy(color),
                  radii=npy(radii))
        g = ref.decode_geometry(geom, P, settings)
        vis = npy(radii) > 0
        for k in ("depths", "means2D", "rects2D", "conic_opacity", "rgb", "tiles_touched", "clamped"):
            v = npy(g[k]).copy()
            v[~vis] = 0  # rows of culled Gaussians are never written by the reference (stale memory)
            fx["geom_" + k] = v
        b = ref.decode_binning(binning, R)
        fx["point_list"] = npy(b["point_list"])
        fx["point_list_keys"] = npy(b["point_list_keys"])
        im = ref.decode_image(img, W, H)
        fx["ranges"] = npy(im["ranges"])
        fx["final_T"] = npy(im["final_T"])
        if settings["sort_settings"]["sort_mode"] != 3:  # HIER leaves n_contrib unwritten
            fx["n_contrib"] = npy(im["n_contrib"])
        if bwd:
            dL = S.make_upstream_grad(W, H, seed + 1000).to(dev)
            names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
                     "dL_drot"]
            g1 = ref.backward(sc, cam, settings, out, dL)
            g2 = ref.backward(sc, cam, settings, out, dL)
            torch.cuda.synchronize()
            for n, x, y in zip(names, g1, g2):
                fx[n] = npy(x)
                # run-to-run noise floor of the reference's atomics, relative to max|grad|
                fx[n + "_noise"] = float((x.double() - y.double()).abs().max() / x.double().abs().max().clamp_min(1e-30))
        path = os.path.join(a.out, name + ".npz")
        np.savez_compressed(path, **fx)
        lens = im["ranges"][:, 1] - im["ranges"][:, 0]
        print(f"{name}: R={R} visible={(radii > 0).sum().item()} max_list={int(lens.max())} "
              f"-> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
