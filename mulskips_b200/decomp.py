"""x-decomposition helpers for N GPUs of one box: the global periodic box is cut into N blocks
along x on super-tile (48-unit) boundaries, one process per GPU; every rank holds its block plus one
24-unit halo tile column of each neighbour.  Mirrors the arithmetic of mskc_create_slab / halo_exchange
(csrc/mskc_api.cu, csrc/mskc_parallel.cuh) for host-side code and tests."""
from __future__ import annotations

HALO = 24      # one tile column
UNIT = 48      # 2x2-coloured 24-unit tiles


def slab_block(rank: int, world: int, global_lenx: int):
    """(x0, owned, local_lenx): first global x of the rank's block, its width, and the local array width."""
    if global_lenx % (UNIT * world):
        raise ValueError(f"global LenX {global_lenx} does not split into {world} blocks on {UNIT}-unit boundaries")
    owned = global_lenx // world
    if owned < UNIT:
        raise ValueError("fewer than two tile columns per rank")
    return rank * owned, owned, owned + 2 * HALO


def neighbours(rank: int, world: int):
    """(left, right) ranks of the periodic 1-D process grid (a single rank is its own neighbour)."""
    return (rank - 1) % world, (rank + 1) % world


def to_local(x_global: int, rank: int, world: int, global_lenx: int) -> int:
    """local x of a global x that lies in the rank's block or in one of its halo columns"""
    x0, owned, _ = slab_block(rank, world, global_lenx)
    d = (x_global - x0) % global_lenx
    if d >= global_lenx - HALO:
        d -= global_lenx
    if not -HALO <= d < owned + HALO:
        raise ValueError("x is outside the rank's block and halos")
    return d + HALO


def seam_shift(owned: int, side: int) -> int:
    """index shift that maps a cell of my left (side 0) / right (side 1) 48-unit strip onto the neighbour's copy"""
    return owned if side == 0 else -owned


def aggregate(dist, values, ops, device="cuda"):
    """all-reduce a few scalars: ops[i] in {'max','sum'}"""
    import torch
    out = []
    for v, op in zip(values, ops):
        t = torch.tensor([float(v)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        out.append(t.item())
    return out
