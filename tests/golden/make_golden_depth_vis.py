#!/usr/bin/env python
"""Golden render_depth images: the UNMODIFIED reference build (oracle/_ref) run with render_depth=True
(DebugVisualization::Depth, rasterize_points.cu:104-107 -- the one debug visualisation its Python API reaches) on the
committed golden scenes, one image per sort mode / configuration.  They pin the oracle's debug-visualisation
accumulators (stp_oracle.c: vis_accum / vis_store, oracle/cpu_oracle.py: colormap) on the CPU; the sort-error
visualisations share those accumulators and depths and differ only in the quantity summed.

    gpurun -- python tests/golden/make_golden_depth_vis.py        # on the B200 box; writes gpurun_out/golden/depth_vis.npz
    cp gpurun_out/golden/depth_vis.npz tests/golden/              # here, then commit
"""
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

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["global_default", "global_distance_ewa", "global_tbc_ptdmax", "hier_default", "hier_preset", "hier_q16_20",
         "hier_sparse", "hier_long", "kbuffer16", "kbuffer4", "full_sort", "full_sort_long"]


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    fx = {}
    for name in CASES:
        case = np.load(os.path.join(HERE, name + ".npz"))
        s = np.load(os.path.join(HERE, "scene_" + str(case["scene"]) + ".npz"))
        settings, deg = json.loads(str(case["settings"])), int(case["sh_degree"])
        t = lambda k: torch.from_numpy(np.ascontiguousarray(s[k]))  # noqa: E731
        sc = S.Scene(t("means3D"), t("scales"), t("rotations"), t("opacities"),
                     t("shs")[:, :(deg + 1) ** 2].contiguous(), deg)
        cam = S.Camera(int(s["H"]), int(s["W"]), float(s["tanfovx"]), float(s["tanfovy"]), t("viewmatrix"), t("projmatrix"),
                       t("inv_viewprojmatrix"), t("campos"), t("bg"))
        sc, cam = S.to_device(sc, dev), S.to_device(cam, dev)
        out = ref.forward(sc, cam, settings, render_depth=True)
        torch.cuda.synchronize()
        assert out[0] == int(case["R"]), name
        fx[name] = out[1].detach().cpu().numpy()
        print(name, fx[name].shape, float(fx[name].min()), float(fx[name].max()))
    np.savez_compressed(os.path.join(out_dir, "depth_vis.npz"), **fx)


if __name__ == "__main__":
    main()
