import sys, json, numpy as np, time
sys.path.insert(0, '/root/repo')
from oracle import cpu_oracle as co
G = '/root/repo/tests/golden/'
def run(name):
    fx = np.load(G + name + '.npz')
    sc = np.load(G + 'scene_' + str(fx['scene']) + '.npz')
    st = json.loads(str(fx['settings']))
    deg = int(fx['sh_degree']); M = (deg + 1) ** 2
    t = time.time()
    o = co.Oracle(st, sc['means3D'], sc['scales'], sc['rotations'], sc['opacities'], sc['shs'][:, :M], deg, sc['viewmatrix'],
                  sc['projmatrix'], sc['inv_viewprojmatrix'], sc['campos'], sc['bg'], float(sc['tanfovx']), float(sc['tanfovy']),
                  int(sc['W']), int(sc['H']))
    dt = time.time() - t
    vis = fx['radii'] > 0
    msg = [f"{name:22s} R {o.R}/{int(fx['R'])} radii!= {(o.radii != fx['radii']).sum()}"]
    if o.R == int(fx['R']):
        msg.append(f"plist!= {(o.point_list != fx['point_list']).sum()} ranges!= {(o.ranges != fx['ranges']).sum()}")
    for k in ('depths', 'means2D', 'conic_opacity'):
        a, b = getattr(o, k)[vis], fx['geom_' + k][vis]
        msg.append(f"{k}!= {(a.view(np.int32) != b.view(np.int32)).sum()}")
    d = np.abs(o.out_color - fx['out_color'])
    msg.append(f"img max {d.max():.2e} n>1e-5 {(d > 1e-5).sum()} T max {np.abs(o.final_T - fx['final_T']).max():.2e}")
    if 'n_contrib' in fx: msg.append(f"ncontrib!= {(o.n_contrib != fx['n_contrib']).sum()}")
    if 'dL_dmeans3D' in fx:
        t = time.time()
        g = o.backward(sc['dL_dout'], fx['out_color'])
        for k in ('dL_dmeans2D', 'dL_dcolors', 'dL_dopacity', 'dL_dmeans3D', 'dL_dcov3D', 'dL_dsh', 'dL_dscales', 'dL_drot'):
            a, b = g[k].reshape(-1), fx[k].reshape(-1)
            msg.append(f"{k[3:]} {np.abs(a - b).max() / max(np.abs(b).max(), 1e-30):.1e}")
    print(' | '.join(msg), f"[{dt:.2f}s]")
for n in sys.argv[1:]:
    run(n)
