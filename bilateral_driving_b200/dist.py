"""Multi-GPU sharding of the render path (new: the reference is single-GPU).

The unit of work is a tile row of one camera.  All C*tile_rows rows of the rig are laid end to end
and rank r of G renders the contiguous band ``[R*r/G, R*(r+1)/G)`` - camera sharding when G divides
the number of cameras, tile sharding otherwise (8 GPUs, 6 cameras).  Full-resolution guidance needs
no halo.  Every rank holds a full parameter replica; after the local backward ONE all-reduce (sum)
over a flat fp32 buffer [Gaussian grads | grid grads] gives every rank the single-GPU gradient.
"""
from typing import List, Sequence, Tuple

import torch

from ._lib import TILE


def band_for_rank(rank: int, world: int, n_cams: int, height: int) -> Tuple[int, int]:
    tile_h = (height + TILE - 1) // TILE
    total = n_cams * tile_h
    return (total * rank) // world, (total * (rank + 1)) // world


def band_pixel_rows(rb: int, re: int, n_cams: int, height: int) -> Tuple[int, int]:
    tile_h = (height + TILE - 1) // TILE

    def row_of(gr):
        c, ty = divmod(gr, tile_h)
        return c * height + min(ty * TILE, height)

    return row_of(rb), (row_of(re) if re < n_cams * tile_h else n_cams * height)


def cameras_in_band(rb: int, re: int, height: int) -> List[int]:
    tile_h = (height + TILE - 1) // TILE
    if re <= rb:
        return []
    return list(range(rb // tile_h, (re - 1) // tile_h + 1))


def _covered_by(flat: torch.Tensor, t: torch.Tensor) -> bool:
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
    return t.is_contiguous() and lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= hi


def allreduce_grads(tensors: Sequence[torch.Tensor], group=None, flat: torch.Tensor = None) -> None:
    """One all-reduce (SUM) of every gradient: a single coalesced NCCL group launch in place on GPUs, a
    single flat buffer elsewhere (gloo).  No-op for a single process."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [t for t in tensors if t is not None]
    if flat is not None:
        # the render backward wrote the Gaussian gradients as views of ONE flat buffer: reduce it in place
        rest = [g for g in grads if not _covered_by(flat, g)]
        if len(rest) < len(grads):
            grads = [flat] + rest
    backend = dist.get_backend(group)
    if backend == "nccl" and hasattr(dist, "_coalescing_manager"):
        # one fused NCCL launch (ncclGroupStart/End) over the gradient tensors in place: no staging copy
        with dist._coalescing_manager(group=group, device=grads[0].device, async_ops=False):
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
