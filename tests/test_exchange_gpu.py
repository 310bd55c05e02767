"""GPU: the splat-space gradient exchange (dist.py / RenderCfg.exchange_group) against the single-process render.
Two ranks share cuda:0 (gloo moves the records through the host - this checks the arithmetic, not the transport):
each renders its band of tile rows, all-gathers the per-splat gradient records in the backward and runs the
projection backward over both ranks' records.  Every rank must end with the gradient of the FULL render."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
SIZES = ((4, 4, 2), (8, 8, 4), (6, 5, 3))


def _inputs():
    from bilateral_driving_b200 import synthetic as S
    from oracle.make_golden import small_scene

    p, vm, Ks, W, H = small_scene(torch.float32)
    grids = S.make_grids(2, SIZES)
    sky, _ = S.make_images(2, H, W)
    return p, vm, Ks, W, H, grids, sky


def _run(rank, world, group, mode="splats"):
    from bilateral_driving_b200.dist import allreduce_grads, band_for_rank
    from bilateral_driving_b200.render import render_fused

    p, vm, Ks, W, H, grids, sky = _inputs()
    Cn = 2
    rb, re = band_for_rank(rank, world, Cn, H) if world > 1 else (0, -1)
    c_p = {k: v.cuda().requires_grad_(True) for k, v in p.items()}
    c_g = [g.cuda().requires_grad_(True) for g in grids]
    out = render_fused(c_p, vm.cuda(), Ks.cuda(), W, H, sky=None, grid_slots=[[g[c] for g in c_g] for c in range(Cn)],
                       bil_sizes=SIZES, near_plane=0.1, row_begin=rb, row_end=re, dense_info=False,
                       exchange_group=group, exchange_mode=mode)
    r0, r1 = out["pixel_rows"]
    gen = torch.Generator().manual_seed(4)
    G = torch.randn(Cn * H, W, 3, generator=gen).cuda()
    ((out["rgb"] * G[r0:r1]).sum() + 0.1 * out["depth"].sum()).backward()
    if world > 1:
        assert out["info"].get("grads_are_global")
        if not out["info"].get("grids_are_global"):   # "compact" reduces the grid-slot gradients in the same all-reduce
            allreduce_grads([g.grad for g in c_g], group=None)
    return {k: v.grad.cpu() for k, v in c_p.items()}, [g.grad.cpu() for g in c_g]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = _run(rank, world, True, mode)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["splats", "compact"])
def test_exchange_of_splat_records_gives_every_rank_the_full_gradient(mode):
    """mode "splats": all-gather of the per-splat records; mode "compact": all-reduce with the SH gradient as one colour
    cotangent per (camera, Gaussian) + bds_sh_expand_bwd."""
    full_p, full_g = _run(0, 1, None)
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, mode), nprocs=world, join=True)
    for r in range(world):
        part_p, part_g = ret[r]
        for k in full_p:
            err = float((part_p[k] - full_p[k]).abs().max() / full_p[k].abs().max().clamp(min=1e-12))
            assert err < 1e-4, (r, k, err)
        for a, b in zip(part_g, full_g):
            assert float((a - b).abs().max() / b.abs().max().clamp(min=1e-12)) < 1e-4, r
