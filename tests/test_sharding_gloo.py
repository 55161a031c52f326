"""N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1 rendezvous) covering the band partition,
the image all-gather and the parameter-gradient all-reduce of stp_sharding.py; plus a CPU-oracle check that
band-local screen-space gradients sum to the full ones (the property the multi-GPU backward relies on)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import stp_sharding as sh


def test_equal_and_balanced_bands_partition_all_rows():
    for gy in (1, 5, 68, 135):
        for world in (1, 2, 4, 8):
            b = sh.equal_bands(gy, world)
            assert b[0][0] == 0 and b[-1][1] == gy and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    w = [0, 0, 10, 10, 10, 10, 0, 40]
    b = sh.balanced_bands(w, 2)
    assert b == [(0, 6), (6, 8)]  # 40 | 40
    b4 = sh.balanced_bands(w, 4)
    assert b4[0][0] == 0 and b4[-1][1] == 8 and all(b4[i][1] == b4[i + 1][0] for i in range(3))
    assert sh.balanced_bands([0, 0, 0], 2) == sh.equal_bands(3, 2)
    r = torch.tensor([[0, 3], [3, 5], [5, 5], [5, 9]], dtype=torch.int32)
    assert sh.row_weights_from_ranges(r, 2, 2).tolist() == [5, 4]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W, P, M = 72, 100, 7, 4
        gy = (H + 15) // 16
        bands = sh.equal_bands(gy, world)
        # every rank "renders" its band of a known image
        truth = torch.arange(3 * H * W, dtype=torch.float32).view(3, H, W)
        local = torch.full((3, H, W), -1.0)
        s, e = bands[rank][0] * 16, min(bands[rank][1] * 16, H)
        local[:, s:e] = truth[:, s:e]
        full = sh.gather_image_bands(local, bands)
        ok_img = bool(torch.equal(full, truth))
        # parameter-gradient slab: sum over ranks, views alias the slab
        n = sh.param_slab_numel(P, M)
        slab = torch.full((n,), float(rank + 1))
        sh.all_reduce_param_grads(slab)
        m3, shg, op, sc, ro = sh.split_param_slab(slab, P, M)
        ok_slab = bool((slab == sum(range(1, world + 1))).all()) and m3.shape == (P, 3) and shg.shape == (P, M, 3) \
            and op.shape == (P, 1) and sc.shape == (P, 3) and ro.shape == (P, 4) and ro.data_ptr() > m3.data_ptr()
        q.put((rank, ok_img, ok_slab))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_image_gather_and_grad_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_band_gradients_sum_to_full(golden):
    """CPU oracle: zeroing dL_dpix outside a band gives band-local gradients; the bands sum to the full gradient
    (linearity of the backward pass in dL_dpix) -- what tile-band sharding relies on."""
    f = golden("global_default")
    o = f.oracle()
    dL = f.scene["dL_dout"]
    full = o.backward(dL, f.fx["out_color"])
    cut = 32
    a, b = dL.copy(), dL.copy()
    a[:, cut:] = 0
    b[:, :cut] = 0
    ga, gb = o.backward(a, f.fx["out_color"]), o.backward(b, f.fx["out_color"])
    for k in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drot"):
        s = ga[k] + gb[k]
        m = np.abs(full[k]).max()
        assert np.abs(s - full[k]).max() <= 2e-5 * m, k
