"""pytest configuration: the `gpu` marker, import paths and the golden-fixture loader.

`-m "not gpu"` : oracle vs. golden vectors, host logic, C-ABI symbol export (runs without a GPU).
`-m gpu`       : parity tests proper -- the CUDA path called through the C ABI (ctypes) against the
                 golden fixtures (outputs of the unmodified reference build), the CPU oracle and, when
                 oracle/_ref is present, the reference build itself on the same device.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stopthepop-rasterization_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a machine without a GPU is a skip, never a silent pass on some fallback
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drot"]

CASES = ["global_default", "global_distance_ewa", "global_tbc_ptdmax", "global_ptdcenter_lb", "hier_default",
         "hier_preset", "hier_cull_only", "hier_q16_20", "hier_q8_12", "hier_sparse", "hier_long", "kbuffer16",
         "kbuffer4", "full_sort", "full_sort_long", "c1_config"]


class Fixture:
    """one golden case: inputs of its scene + every output the reference build produced."""

    def __init__(self, name):
        self.name = name
        self.fx = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.scene = np.load(os.path.join(GOLDEN, "scene_" + str(self.fx["scene"]) + ".npz"))
        self.settings = json.loads(str(self.fx["settings"]))
        self.deg = int(self.fx["sh_degree"])
        self.M = (self.deg + 1) ** 2
        self.W, self.H, self.P = int(self.scene["W"]), int(self.scene["H"]), int(self.scene["P"])
        self.has_bwd = "dL_dmeans3D" in self.fx
        self.sort_mode = self.settings["sort_settings"]["sort_mode"]
        self.hier_cull = self.sort_mode == 3 and self.settings["culling_settings"]["hierarchical_4x4_culling"]

    def shs(self):
        return np.ascontiguousarray(self.scene["shs"][:, :self.M])

    def oracle(self):
        from oracle import cpu_oracle as co
        s = self.scene
        return co.Oracle(self.settings, s["means3D"], s["scales"], s["rotations"], s["opacities"], self.shs(), self.deg,
                         s["viewmatrix"], s["projmatrix"], s["inv_viewprojmatrix"], s["campos"], s["bg"],
                         float(s["tanfovx"]), float(s["tanfovy"]), self.W, self.H)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Fixture(name)
        return cache[name]
    return get
