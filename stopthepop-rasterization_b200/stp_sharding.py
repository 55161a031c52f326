"""Multi-GPU sharding of the rasterizer hot path: one process per GPU, torch.distributed.

Two ways to shard, one exchange step each:

* views   -- replicated Gaussians, one camera per rank; exchange = ONE all-reduce(sum) of the contiguous
             parameter-gradient slab returned by `_C.rasterize_gaussians_backward(..., want_param_slab=True)`.
* bands   -- one view, contiguous bands of 16-pixel tile rows per rank (`tile_band=(row0,row1)`): every rank
             preprocesses all Gaussians but bins / sorts / renders only its band, so concatenating the bands
             reproduces the single-GPU point_list / ranges / image bit for bit.  Exchange = ONE all-gather of the
             image bands (forward) and ONE all-reduce(sum) of the parameter-gradient slab (backward; the
             per-Gaussian backward is linear in the band-local screen-space gradients).

The collectives are torch.distributed calls (NCCL over NVLink on the GPU box, gloo in the CPU tests); nothing here
touches the kernels.
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def equal_bands(grid_y: int, world: int) -> List[Tuple[int, int]]:
    """contiguous tile-row bands of (almost) equal height; empty bands when world > grid_y."""
    edges = [(grid_y * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def balanced_bands(row_weights: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """contiguous bands whose summed weight (e.g. instances per tile row of the previous frame) is as equal as a
    greedy prefix split allows: band r ends at the first row where the running sum reaches (r+1)/world of the total."""
    w = torch.as_tensor(row_weights, dtype=torch.float64)
    grid_y = int(w.numel())
    total = float(w.sum())
    if total <= 0 or world == 1:
        return equal_bands(grid_y, world)
    csum = torch.cumsum(w, 0)
    edges = [0]
    for r in range(1, world):
        target = total * r / world
        e = int(torch.searchsorted(csum, torch.tensor(target, dtype=torch.float64)).item()) + 1
        edges.append(min(max(e, edges[-1]), grid_y))
    edges.append(grid_y)
    return [(edges[r], edges[r + 1]) for r in range(world)]


def row_weights_from_ranges(ranges: torch.Tensor, grid_x: int, grid_y: int) -> torch.Tensor:
    """instances per tile row from the image arena's `ranges[tiles,2]` (view_image(...)['ranges'])."""
    lens = (ranges[:, 1] - ranges[:, 0]).to(torch.int64).view(grid_y, grid_x)
    return lens.sum(1)


def all_reduce_param_grads(slab: torch.Tensor, group=None) -> torch.Tensor:
    """the one gradient exchange of view sharding: in-place all-reduce(sum) of the flat parameter-gradient slab of
    _C.rasterize_gaussians_backward(want_param_slab=True).  Slab order: [sh(3MP) | means3D(3P) | scales(3P) | rot(4P) |
    opacity(P)], every sub-array starting on a multiple of 4 floats -- slice it with split_param_slab, never by hand."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(slab, op=dist.ReduceOp.SUM, group=group)
    return slab


def split_param_slab(slab: torch.Tensor, P: int, M: int):
    """views into the slab: (dL_dmeans3D[P,3], dL_dsh[P,M,3], dL_dopacity[P,1], dL_dscales[P,3], dL_drot[P,4])."""
    widths = [3 * M, 3, 3, 4, 1]  # slab order of _C.rasterize_gaussians_backward: sh means3D scales rot opacity,
    out, off = [], 0              # every sub-array starting on a multiple of 4 floats (_C.slab_offsets)
    for w in widths:
        off = (off + 3) // 4 * 4
        out.append(slab[off:off + w * P])
        off += w * P
    return out[1].view(P, 3), out[0].view(P, M, 3), out[4].view(P, 1), out[2].view(P, 3), out[3].view(P, 4)


def param_slab_numel(P: int, M: int) -> int:
    """length (floats) of the parameter-gradient slab for P Gaussians with M SH coefficients"""
    off = 0
    for w in (3 * M, 3, 3, 4, 1):
        off = (off + 3) // 4 * 4 + w * P
    return (off + 3) // 4 * 4


def gather_image_bands(local_img: torch.Tensor, bands: Sequence[Tuple[int, int]], group=None) -> torch.Tensor:
    """assemble the full [C,H,W] image from the per-rank band renders.  `local_img` is this rank's full-size
    output (only the rows of its band are meaningful).  ONE all-gather of equally padded band slabs."""
    C, H, W = local_img.shape
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_img
    rank = dist.get_rank(group)
    px = [(min(b0 * 16, H), min(b1 * 16, H)) for b0, b1 in bands]
    max_rows = max(1, max(e - s for s, e in px))
    send = local_img.new_zeros((C, max_rows, W))
    s, e = px[rank]
    send[:, :e - s] = local_img[:, s:e]
    recv = local_img.new_empty((world, C, max_rows, W))
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group) if local_img.is_cuda else \
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    full = torch.empty_like(local_img)
    for r, (s, e) in enumerate(px):
        full[:, s:e] = recv[r, :, :e - s]
    return full
