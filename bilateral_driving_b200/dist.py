"""Multi-GPU sharding of the render path (new: the reference is single-GPU).

The unit of work is a tile row of one camera.  All C*tile_rows rows of the rig are laid end to end
and rank r of G renders the contiguous band ``[R*r/G, R*(r+1)/G)`` - camera sharding when G divides
the number of cameras, tile sharding otherwise (8 GPUs, 6 cameras).  Full-resolution guidance needs
no halo.  Every rank holds a full parameter replica; after the local backward ONE all-reduce (sum)
over a flat fp32 buffer [Gaussian grads | grid grads] gives every rank the single-GPU gradient.
"""
from typing import List, Optional, Sequence, Tuple

import torch

from ._lib import TILE


def band_for_rank(rank: int, world: int, n_cams: int, height: int) -> Tuple[int, int]:
    tile_h = (height + TILE - 1) // TILE
    total = n_cams * tile_h
    return (total * rank) // world, (total * (rank + 1)) // world


def balanced_bands(row_weights: Sequence[float], world: int, rows_per_camera: Optional[int] = None,
                   camera_cost: float = 0.0) -> List[Tuple[int, int]]:
    """Contiguous bands of the global tile rows with about equal COST instead of equal row counts.  The work of a
    tile row is far from uniform (rows at the horizon of a driving scene hold most of the records), and with equal row
    counts every step waits for the heaviest band at the gradient exchange.  ``row_weights[i]`` = cost estimate of
    global tile row i (e.g. records of the previous step + a per-tile constant); a band additionally pays
    ``camera_cost`` for every camera it touches (projection and emission run once per camera of the band), cameras
    being runs of ``rows_per_camera`` rows.  Min-max partition (binary search on the largest band cost, greedy
    fill); returns ``world`` bands ``[begin, end)`` that partition the rows, every band non-empty when there are at
    least ``world`` rows."""
    n = len(row_weights)
    w = [max(float(x), 0.0) for x in row_weights]
    world = max(1, min(world, n)) if n else world

    def cams(b, e):
        if not rows_per_camera or e <= b:
            return 0
        return (e - 1) // rows_per_camera - b // rows_per_camera + 1

    def fill(limit):
        """Greedy: longest bands whose cost stays <= limit, keeping one row for every band still to come."""
        cuts, b = [0], 0
        for r in range(world):
            e, acc = b, 0.0
            last = n - (world - 1 - r)               # rows this band may take at most
            while e < last:
                nxt = acc + w[e] + camera_cost * (cams(b, e + 1) - cams(b, e))
                if e > b and nxt > limit:
                    break
                acc, e = nxt, e + 1
            cuts.append(e)
            b = e
        return cuts

    lo = max(w) if w else 0.0
    hi = sum(w) + camera_cost * (cams(0, n) + world) + 1.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if fill(mid)[-1] >= n:
            hi = mid
        else:
            lo = mid
    cuts = fill(hi)
    cuts[-1] = n
    for r in range(world - 1, 0, -1):                # degenerate weights: keep every band non-empty
        cuts[r] = min(cuts[r], cuts[r + 1] - 1)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


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


# ---- gradient exchange in SPLAT space ---------------------------------------------------------------------------------
# The dense parameter gradient of the Gaussians is 236 B x N per rank (472 MB at 2 M Gaussians) and its all-reduce is
# the fixed cost that bounds strong scaling.  What a rank's backward actually produces is far smaller: one 48-byte
# gradient record (the pixel moments of composite_bwd) per splat VISIBLE in its band, next to the 48-byte forward
# record of that splat.  Every parameter gradient is a linear function of those records (project_bwd), so instead of
# reducing 472 MB the ranks ALL-GATHER the records (96 B x visible splats, about 4 x less data and one direction
# instead of reduce-scatter + all-gather) and each rank runs the projection backward over all of them.  A splat that
# reaches two bands simply appears twice, with partial moments: the projection backward is linear in them.
def gather_splat_counts(n_local: int, device, group=None) -> Tuple[List[int], torch.Tensor]:
    """All ranks' record counts: a host list (sizes the gather buffers) and the device total (int32[1], what
    bds_project_bwd reads as counters[0])."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mine = torch.tensor([n_local], device=device, dtype=torch.int64)
    every = torch.empty(world, device=device, dtype=torch.int64)
    dist.all_gather_into_tensor(every, mine, group=group)
    return [int(v) for v in every.tolist()], every.sum().to(torch.int32).reshape(1)


def allgather_rows(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Concatenation over ranks of the first counts[r] rows of every rank's ``local`` ([>= counts[rank], K])."""
    import torch.distributed as dist

    rank = dist.get_rank(group)
    K = local.shape[1]
    total = int(sum(counts))
    out = torch.empty(max(total, 1), K, device=local.device, dtype=local.dtype)
    mine = local[:counts[rank]]
    if dist.get_backend(group) == "nccl":
        # uneven all-gather: NCCL runs it as one grouped launch, every rank writes straight into its slice
        pieces = list(out[:total].split(list(counts)))
        dist.all_gather(pieces, mine.contiguous(), group=group)
        return out
    # portable path (gloo: equal sizes only): pad to the longest piece, then compact
    m = max(max(counts), 1)
    padded = torch.zeros(m, K, device=local.device, dtype=local.dtype)
    padded[:counts[rank]] = mine
    every = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(every, padded, group=group)
    off = 0
    for r, c in enumerate(counts):
        out[off:off + c] = every[r][:c]
        off += c
    return out
