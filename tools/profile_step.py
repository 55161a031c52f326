#!/usr/bin/env python
"""One fwd+bwd step of a bench workload between cudaProfilerStart/Stop, for ncu --profile-from-start off.
usage: python tools/profile_step.py <workload> [warm-up steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stopthepop-rasterization_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import stp_scenes as S  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3b"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene_name, overrides, desc = bench.WORKLOADS[wl]
settings = S.default_settings_dict(**overrides)
cid, P, W, H = S.CONFIGS[scene_name]
dev = torch.device("cuda:0")
sc, cam = S.make_config(scene_name)
sc, cam = S.to_device(sc, dev), S.to_device(cam, dev)
dL = S.make_upstream_grad(W, H, 2000 + cid).to(dev)
e = torch.empty(0, device=dev)


def step():
    out = _C.rasterize_gaussians(cam.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e, cam.viewmatrix,
                                 cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx, cam.tanfovy, H, W, sc.shs, 3, cam.campos,
                                 False, settings, False, False)
    _C.rasterize_gaussians_backward(cam.bg, sc.means3D, out[2], sc.opacities, e, sc.scales, sc.rotations, 1.0, e,
                                    cam.viewmatrix, cam.projmatrix, cam.inv_viewprojmatrix, cam.tanfovx, cam.tanfovy, out[1],
                                    dL, sc.shs, 3, cam.campos, out[3], out[0], out[4], out[5], settings, False)


for _ in range(warm):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(wl, "profiled one step")
