import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stopthepop-rasterization_b200"))
import numpy as np, torch
from conftest import Fixture
from test_gpu_parity import run_ours, npy
f = Fixture(sys.argv[1] if len(sys.argv) > 1 else "full_sort")
r = run_ours(f, backward=False)
fx = f.fx
nc, nr = npy(r["image"]["n_contrib"]), fx["n_contrib"]
T, Tr = npy(r["image"]["final_T"]), fx["final_T"]
d = np.abs(npy(r["out_color"]) - fx["out_color"]).max(0)
bad = np.argwhere((nc != nr) | (d > 1e-5))
print("bad pixels", len(bad), "of", nc.size)
rg = fx["ranges"]
gx = (f.W + 15) // 16
for (y, x) in bad[:25]:
    t = (y // 16) * gx + x // 16
    print(f"px {x},{y} tile {t} len {rg[t,1]-rg[t,0]} ncontrib ours {nc[y,x]} ref {nr[y,x]}  T ours {T[y,x]:.6f} ref {Tr[y,x]:.6f} dcol {d[y,x]:.2e}")
