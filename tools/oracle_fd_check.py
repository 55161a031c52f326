import sys, json, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import cpu_oracle as co
G = '/root/repo/tests/golden/'
name = sys.argv[1]; gid = int(sys.argv[2])
fx = np.load(G + name + '.npz'); sc = np.load(G + 'scene_' + str(fx['scene']) + '.npz')
st = json.loads(str(fx['settings'])); deg = int(fx['sh_degree']); M = (deg + 1) ** 2
def fwd(op):
    o = co.Oracle(st, sc['means3D'], sc['scales'], sc['rotations'], op, sc['shs'][:, :M], deg, sc['viewmatrix'],
              sc['projmatrix'], sc['inv_viewprojmatrix'], sc['campos'], sc['bg'], float(sc['tanfovx']), float(sc['tanfovy']), int(sc['W']), int(sc['H']))
    return o
op = sc['opacities'].copy()
o0 = fwd(op)
g = o0.backward(sc['dL_dout'], fx['out_color'])
for eps in (1e-2, 1e-3):
    a = op.copy(); a[gid] += eps; b = op.copy(); b[gid] -= eps
    fa, fb = fwd(a).out_color.astype(np.float64), fwd(b).out_color.astype(np.float64)
    print("eps", eps, "FD dL/dopacity", ((fa - fb) * sc['dL_dout']).sum() / (2 * eps))
print("oracle", g['dL_dopacity'][gid], "ref", fx['dL_dopacity'][gid])
