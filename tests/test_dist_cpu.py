"""CPU, world_size 2, gloo: the host-side multi-GPU logic (tile-row bands + the single flat gradient
all-reduce) without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bilateral_driving_b200.dist import allreduce_grads, balanced_bands, band_for_rank, band_pixel_rows, cameras_in_band


def test_bands_partition_the_rig():
    H, C = 1080, 6
    tile_h = 68
    for world in (1, 2, 3, 4, 6, 8):
        covered = []
        rows = []
        for r in range(world):
            rb, re = band_for_rank(r, world, C, H)
            covered += list(range(rb, re))
            r0, r1 = band_pixel_rows(rb, re, C, H)
            rows.append((r0, r1))
            cams = cameras_in_band(rb, re, H)
            assert cams == sorted(set(g // tile_h for g in range(rb, re)))
        assert covered == list(range(C * tile_h))
        assert rows[0][0] == 0 and rows[-1][1] == C * H
        for a, b in zip(rows[:-1], rows[1:]):
            assert a[1] == b[0]
    # 6 GPUs, 6 cameras -> one camera each; 8 GPUs -> bands cut cameras (tile sharding)
    assert [cameras_in_band(*band_for_rank(r, 6, C, H), H) for r in range(6)] == [[c] for c in range(6)]
    assert any(len(cameras_in_band(*band_for_rank(r, 8, C, H), H)) == 2 for r in range(8))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a toy "loss" that is a sum over pixel rows: each rank owns its band of rows, gradients add up
        H, C, W = 40, 3, 8
        g = torch.Generator(); g.manual_seed(0)
        w = torch.randn(5, generator=g, requires_grad=True)
        img = torch.randn(C * H, W, generator=g)
        rb, re = band_for_rank(rank, world, C, H)
        r0, r1 = band_pixel_rows(rb, re, C, H)
        loss = (img[r0:r1].sum(1)[:, None] * w[None, :]).sum() / (C * H * W)
        loss.backward()
        grid_grad = torch.full((3,), float(rank + 1))
        allreduce_grads([w.grad, None, grid_grad])
        full = (img.sum(1)[:, None] * torch.ones(1, 5)).sum(0) / (C * H * W)
        ok = torch.allclose(w.grad, full, atol=1e-6) and torch.allclose(grid_grad, torch.full((3,), 3.0))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_band_sharded_gradients_allreduce_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_allreduce_is_noop_without_process_group():
    t = torch.ones(3)
    allreduce_grads([t])
    assert torch.equal(t, torch.ones(3))


def _exchange_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bilateral_driving_b200.dist import allgather_rows, gather_splat_counts

        n_local = [5, 0, 3][rank]
        local = torch.full((8, 12), float(rank)) + torch.arange(8)[:, None] * 0.01   # capacity 8, first n_local rows valid
        counts, total = gather_splat_counts(n_local, torch.device("cpu"))
        out = allgather_rows(local, counts)
        ret[rank] = (counts, int(total), out[:sum(counts)].clone())
    finally:
        dist.destroy_process_group()


def test_splat_record_exchange_gathers_uneven_pieces():
    """dist.gather_splat_counts / allgather_rows (the splat-space gradient exchange) over gloo, world size 3, with an
    empty rank: every rank ends with the concatenation of the valid rows of all ranks, in rank order."""
    world = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        counts, total, out = ret[r]
        assert counts == [5, 0, 3] and total == 8 and out.shape == (8, 12)
        assert torch.allclose(out[:5, 0], 0.0 + torch.arange(5) * 0.01) and torch.allclose(out[5:, 0], 2.0 + torch.arange(3) * 0.01)


def test_balanced_bands_partition_rows_by_weight():
    """dist.balanced_bands: contiguous, exhaustive, non-empty bands whose weights are far closer to each other than an
    equal row split when the rows are not uniform (horizon-heavy driving scenes)."""
    import math

    rows = 6 * 68
    w = [100.0 + 5000.0 * math.exp(-(((i % 68) - 34) / 6.0) ** 2) for i in range(rows)]   # heavy rows at the horizon
    for world in (2, 3, 4, 8):
        bands = balanced_bands(w, world)
        assert bands[0][0] == 0 and bands[-1][1] == rows
        assert all(b[0] < b[1] for b in bands) and all(a[1] == b[0] for a, b in zip(bands[:-1], bands[1:]))
        sums = [sum(w[b[0]:b[1]]) for b in bands]
        equal = [sum(w[rows * r // world:rows * (r + 1) // world]) for r in range(world)]
        assert max(sums) <= max(equal) + 1e-6
        assert max(sums) / (sum(w) / world) < 1.15
    assert balanced_bands([1.0] * 3, 3) == [(0, 1), (1, 2), (2, 3)]
    z = balanced_bands([0.0] * 8, 4)
    assert z[0][0] == 0 and z[-1][1] == 8 and all(b[0] < b[1] for b in z)
    # a per-camera cost: 8 bands over 6 cameras of 68 rows - bands that straddle a camera boundary get fewer rows
    flat = [1000.0] * rows
    plain = balanced_bands(flat, 8)
    costly = balanced_bands(flat, 8, rows_per_camera=68, camera_cost=30000.0)
    assert costly[0][0] == 0 and costly[-1][1] == rows and all(a[1] == b[0] for a, b in zip(costly[:-1], costly[1:]))

    def cost(b):
        return sum(flat[b[0]:b[1]]) + 30000.0 * ((b[1] - 1) // 68 - b[0] // 68 + 1)

    assert max(cost(b) for b in costly) < max(cost(b) for b in plain)
