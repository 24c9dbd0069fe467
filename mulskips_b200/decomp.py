"""x-y decomposition helpers for N GPUs of one box: blocks on super-tile (48-unit)
boundaries, one process per GPU, and the rank aggregation used for whole-job throughput."""
from __future__ import annotations


def slab_grid(world: int):
    """Process grid gx x gy (gx >= gy) as square as possible: 1x1, 2x1, 2x2, 4x2."""
    gy = 1
    for d in range(1, int(world ** 0.5) + 1):
        if world % d == 0:
            gy = d
    return world // gy, gy


def rank_block(rank: int, world: int, lenx: int, leny: int, unit: int = 48):
    """(x0, y0, lx, ly) of the rank's x-y block inside a global lenx x leny lattice."""
    gx, gy = slab_grid(world)
    if lenx % (gx * unit) or leny % (gy * unit):
        raise ValueError(f"global box {lenx}x{leny} does not split into {gx}x{gy} blocks on {unit}-unit boundaries")
    lx, ly = lenx // gx, leny // gy
    return (rank % gx) * lx, (rank // gx) * ly, lx, ly


def aggregate(dist, values, ops, device="cuda"):
    """all-reduce a few scalars: ops[i] in {'max','sum'}"""
    import torch
    out = []
    for v, op in zip(values, ops):
        t = torch.tensor([float(v)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        out.append(t.item())
    return out
