#!/usr/bin/env python
"""Exploratory parity probe (run on the GPU box): ours vs. the reference build, field by field."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stopthepop-rasterization_b200"))
sys.path.insert(0, ROOT)
import torch

import stp_scenes as S
from diff_gaussian_rasterization import _C as ours
from oracle import ref_api as ref


def bits(t):
    return t.contiguous().view(torch.int32)


def cmp_exact(name, a, b, mask=None):
    if mask is not None:
        a, b = a[mask], b[mask]
    if a.dtype.is_floating_point:
        ne = bits(a) != bits(b)
    else:
        ne = a != b
    n = int(ne.sum())
    extra = ""
    if n and a.dtype.is_floating_point:
        d = (a.double() - b.double()).abs()
        rel = d / b.double().abs().clamp_min(1e-30)
        extra = f" maxabs={d.max().item():.3e} maxrel={rel[ne].max().item():.3e}"
    print(f"    {name:18s} mismatches {n:9d} / {a.numel():9d}{extra}")
    return n


def cmp_close(name, a, b):
    d = (a.double() - b.double()).abs()
    scale = b.double().abs().max().item()
    print(f"    {name:18s} max|d|={d.max().item():.3e}  max|ref|={scale:.3e}  rel={d.max().item() / max(scale, 1e-30):.3e}"
          f"  n(|d|>1e-5*max)={(d > 1e-5 * scale).sum().item()}")
    return d.max().item() / max(scale, 1e-30)


def run_case(tag, P, W, H, seed, settings, bwd=True, timing=False):
    print(f"=== {tag}: P={P} {W}x{H} settings={json.dumps(settings['sort_settings'])} cull={json.dumps(settings['culling_settings'])} ewa={settings['proper_ewa_scaling']}")
    sc, cam = S.make_scene(P, W, H, seed)
    dev = torch.device("cuda:0")
    sc, cam = S.to_device(sc, dev), S.to_device(cam, dev)
    e = torch.empty(0, device=dev)
    r = ref.forward(sc, cam, settings)
    o = ours.rasterize_gaussians(cam.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e, cam.viewmatrix,
                                 cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx, cam.tanfovy, H, W, sc.shs,
                                 sc.sh_degree, cam.campos, False, settings, False, False)
    torch.cuda.synchronize()
    print(f"    R ref={r[0]} ours={o[0]}   visible ref={(r[2] > 0).sum().item()} ours={(o[2] > 0).sum().item()}")
    cmp_exact("radii", o[2], r[2])
    vis = r[2] > 0
    rg = ref.decode_geometry(r[3], P, settings)
    og = ours.view_geometry(o[3], P, settings)
    for k in ("depths", "means2D", "rects2D", "conic_opacity", "cov3D", "tiles_touched", "cov3D_inv", "clamped"):
        if k in rg and og.get(k) is not None:
            cmp_exact(k, og[k], rg[k], vis)
    cmp_close("rgb", og["rgb"][vis], rg["rgb"][vis])
    if r[0] == o[0]:
        rb, ob = ref.decode_binning(r[4], r[0]), ours.view_binning(o[4], o[0])
        cmp_exact("point_list", ob["point_list"], rb["point_list"])
        ri, oi = ref.decode_image(r[5], W, H), ours.view_image(o[5], W, H)
        cmp_exact("ranges", oi["ranges"], ri["ranges"])
        cmp_exact("final_T", oi["final_T"], ri["final_T"])
        if settings["sort_settings"]["sort_mode"] != 3:
            cmp_exact("n_contrib", oi["n_contrib"], ri["n_contrib"])
    cmp_exact("out_color(bits)", o[1], r[1])
    cmp_close("out_color", o[1], r[1])
    if bwd:
        dL = S.make_upstream_grad(W, H, seed + 1000).to(dev)
        rgr = ref.backward(sc, cam, settings, r, dL)
        rgr2 = ref.backward(sc, cam, settings, r, dL)
        ogr = ours.rasterize_gaussians_backward(cam.bg, sc.means3D, o[2], sc.opacities, e, sc.scales, sc.rotations, 1.0, e,
                                                cam.viewmatrix, cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx,
                                                cam.tanfovy, o[1], dL, sc.shs, sc.sh_degree, cam.campos, o[3], o[0],
                                                o[4], o[5], settings, False)
        torch.cuda.synchronize()
        names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drot"]
        for n, a, b, b2 in zip(names, ogr, rgr, rgr2):
            floor = (b.double() - b2.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)
            rel = cmp_close(n, a, b)
            print(f"        (reference run-to-run noise floor {floor:.2e})")
    if timing:
        def t_ms(fn, n=5):
            fn(); torch.cuda.synchronize()
            s, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                fn()
            e2.record(); torch.cuda.synchronize()
            return s.elapsed_time(e2) / n
        tr = t_ms(lambda: ref.forward(sc, cam, settings))
        to = t_ms(lambda: ours.rasterize_gaussians(cam.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e,
                                                   cam.viewmatrix, cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx,
                                                   cam.tanfovy, H, W, sc.shs, sc.sh_degree, cam.campos, False, settings,
                                                   False, False))
        print(f"    fwd ms: ref {tr:.3f}  ours {to:.3f}")
        if bwd:
            trb = t_ms(lambda: ref.backward(sc, cam, settings, r, dL))
            tob = t_ms(lambda: ours.rasterize_gaussians_backward(
                cam.bg, sc.means3D, o[2], sc.opacities, e, sc.scales, sc.rotations, 1.0, e, cam.viewmatrix,
                cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx, cam.tanfovy, o[1], dL, sc.shs, sc.sh_degree,
                cam.campos, o[3], o[0], o[4], o[5], settings, False))
            print(f"    bwd ms: ref {trb:.3f}  ours {tob:.3f}")
        o2 = ours.rasterize_gaussians(cam.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e,
                                      cam.viewmatrix, cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx, cam.tanfovy,
                                      H, W, sc.shs, sc.sh_degree, cam.campos, False, settings, False, 2)
        print("    stage ms:", ours.last_timings())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", type=int, default=300000)
    a = ap.parse_args()
    print(torch.cuda.get_device_name(0), "cpu cores", os.cpu_count())
    D = S.default_settings_dict
    run_case("C1 default", 1000, 256, 256, 1001, D())
    run_case("mid default", a.big, 1920, 1080, 1002, D(), timing=True)
    run_case("mid distance", a.big, 1920, 1080, 1002, D(sort_order=1), bwd=False)
    run_case("mid rect+tight", a.big, 1920, 1080, 1002, D(rect_bounding=True, tight_opacity_bounding=True), bwd=False)
    run_case("mid TBC", a.big, 1920, 1080, 1002, D(rect_bounding=True, tight_opacity_bounding=True, tile_based_culling=True), bwd=False)
    run_case("mid TBC+LB", a.big, 1920, 1080, 1002, D(rect_bounding=True, tight_opacity_bounding=True, tile_based_culling=True, load_balancing=True), bwd=False)
    run_case("mid PTD_CENTER", a.big, 1920, 1080, 1002, D(sort_order=2), bwd=False)
    run_case("mid PTD_MAX+TBC", a.big, 1920, 1080, 1002, D(sort_order=3, rect_bounding=True, tight_opacity_bounding=True, tile_based_culling=True), bwd=False)
    run_case("mid EWA", a.big, 1920, 1080, 1002, D(proper_ewa_scaling=True), bwd=True)
    run_case("C2 default", 1000000, 1920, 1080, 1002, D(), timing=True)
