#!/usr/bin/env python
"""Per-pixel blend lists of the REFERENCE backward (one-hot upstream gradients) for one golden case.
Used to characterise the reference's HIER backward under 4x4 culling."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stopthepop-rasterization_b200")); sys.path.insert(0, ROOT)
import stp_scenes as S
from oracle import ref_api as ref
name = sys.argv[1]
G = os.path.join(ROOT, "tests", "golden")
fx = np.load(os.path.join(G, name + ".npz")); scn = np.load(os.path.join(G, f"scene_{fx['scene']}.npz"))
st = json.loads(str(fx["settings"])); deg = int(fx["sh_degree"]); M = (deg + 1) ** 2
dev = torch.device("cuda:0")
t = lambda k: torch.from_numpy(scn[k]).to(dev)
sc = S.Scene(t("means3D"), t("scales"), t("rotations"), t("opacities"), t("shs")[:, :M].contiguous(), deg)
W, H = int(scn["W"]), int(scn["H"])
cam = S.Camera(H, W, float(scn["tanfovx"]), float(scn["tanfovy"]), t("viewmatrix"), t("projmatrix"), t("inv_viewprojmatrix"), t("campos"), t("bg"))
out = ref.forward(sc, cam, st)
rows = []
dL = torch.zeros(3, H, W, device=dev)
for p in range(W * H):
    dL.zero_(); dL[0].view(-1)[p] = 1.0
    g = ref.backward(sc, cam, st, out, dL)
    c = g[1][:, 0]
    nz = torch.nonzero(c).flatten()
    rows.append(np.stack([np.full(len(nz), p), nz.cpu().numpy(), c[nz].cpu().numpy().view(np.int32)], 1))
rows = np.concatenate(rows).astype(np.int64)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"pixel_probe_{name}.npy"), rows)
print(name, rows.shape)
