import sys, json, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import cpu_oracle as co
G = '/root/repo/tests/golden/'
name = sys.argv[1]
fx = np.load(G + name + '.npz'); sc = np.load(G + 'scene_' + str(fx['scene']) + '.npz')
st = json.loads(str(fx['settings'])); deg = int(fx['sh_degree']); M = (deg + 1) ** 2
W, H = int(sc['W']), int(sc['H'])
o = co.Oracle(st, sc['means3D'], sc['scales'], sc['rotations'], sc['opacities'], sc['shs'][:, :M], deg, sc['viewmatrix'],
              sc['projmatrix'], sc['inv_viewprojmatrix'], sc['campos'], sc['bg'], float(sc['tanfovx']), float(sc['tanfovy']), W, H)
rows = np.load(f'/root/repo/gpurun_out/pixel_probe_{name}.npy')
ref = {}
for p, g, v in rows:
    ref.setdefault(int(p), {})[int(g)] = np.int32(v).view(np.float32)
ids = (ctypes.c_int * 4096)(); w = (ctypes.c_float * 4096)()
L = co.lib()
bad = []
for p in range(W * H):
    n = L.orc_debug_hier_pixel(ctypes.byref(o.inp), ctypes.byref(o.settings), o.st, p % W, p // W, ids, w, 4096)
    mine = {}
    for k in range(n): mine[ids[k]] = mine.get(ids[k], 0.0) + w[k]
    r = ref.get(p, {})
    keys = set(mine) | set(r)
    err = max([abs(mine.get(k, 0.0) - float(r.get(k, 0.0))) for k in keys] or [0.0])
    if err > 1e-5: bad.append((p, err, len(mine), len(r)))
print(name, "bad pixels", len(bad), "of", W * H)
for b in bad[:12]:
    p = b[0]; print("  pixel", p % W, p // W, "err %.3e n_mine %d n_ref %d" % b[1:])
if bad:
    p = bad[0][0]
    n = L.orc_debug_hier_pixel(ctypes.byref(o.inp), ctypes.byref(o.settings), o.st, p % W, p // W, ids, w, 4096)
    print("  mine order:", [(ids[k], round(w[k], 5)) for k in range(n)])
    print("  ref  set  :", sorted([(k, round(float(v), 5)) for k, v in ref[p].items()], key=lambda t: -t[1]))
from collections import Counter
pat = Counter(); quad = Counter()
for b in bad[:3187]:
    p = b[0]
    n = L.orc_debug_hier_pixel(ctypes.byref(o.inp), ctypes.byref(o.settings), o.st, p % W, p // W, ids, w, 4096)
    order = [ids[k] for k in range(n)]
    miss = [i for i, g in enumerate(order) if g not in ref.get(p, {})]
    extra = [g for g in ref.get(p, {}) if g not in order]
    pat[(len(miss), len(extra), tuple(np.diff(miss)) if len(miss) <= 8 else 'many', n - (miss[-1] if miss else 0))] += 1
for k, v in pat.most_common(15): print(v, k)
