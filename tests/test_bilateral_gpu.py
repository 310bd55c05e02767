"""GPU parity of the stand-alone bilateral kernels (through the C ABI) against (a) golden vectors
minted from the UNMODIFIED reference and (b) the CPU oracle on seeded inputs.
Tolerances: outputs 1e-5 abs, gradients 1e-3 rel (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


@pytest.mark.parametrize("tag,gf", [("f442", (4, 4, 2)), ("none", None)])
def test_multiscale_golden(golden_dir, tag, gf):
    from bilateral_driving_b200.bilateral import multiscale_bilateral

    d = np.load(os.path.join(golden_dir, "bilateral_ms.npz"))
    idx = int(d["idx"])
    sizes = d["sizes"].tolist()
    rgb = torch.from_numpy(d["rgb"]).cuda().requires_grad_(True)
    G = torch.from_numpy(d["G"]).cuda()
    full = [torch.from_numpy(d[f"grids{i}"]).cuda().requires_grad_(True) for i in range(3)]
    y, affs = multiscale_bilateral(rgb, [g[idx] for g in full], sizes, gf, return_affine=True)
    (y * G).sum().backward()
    assert np.abs(y.detach().cpu().numpy() - d[f"{tag}_out"]).max() < 1e-5
    assert np.abs(affs[2].detach().cpu().numpy().reshape(*rgb.shape[:2], 12) - d[f"{tag}_aff2"]).max() < 1e-5
    assert _rel(rgb.grad.cpu(), torch.from_numpy(d[f"{tag}_vrgb"])) < 1e-3
    for i in range(3):
        assert _rel(full[i].grad.cpu(), torch.from_numpy(d[f"{tag}_vgrid{i}"])) < 1e-3
    # fused (no affine fields) path gives the same result
    y2 = multiscale_bilateral(rgb.detach(), [g[idx].detach() for g in full], sizes, gf)
    assert (y2 - y.detach()).abs().max() == 0


def test_config1_golden(golden_dir):
    from bilateral_driving_b200.bilateral import BilateralAffineTransform

    d = np.load(os.path.join(golden_dir, "bilateral_cfg1.npz"))
    g0 = torch.Generator(); g0.manual_seed(0)
    g2 = torch.Generator(); g2.manual_seed(2)
    rgb = torch.rand(256, 256, 3, generator=g0).cuda().requires_grad_(True)
    G = torch.randn(256, 256, 3, generator=g2).cuda()
    m = BilateralAffineTransform("Affine", n=1, grid_X=16, grid_Y=16, grid_W=8).cuda()
    with torch.no_grad():
        m.bil_grids.grids.copy_(torch.from_numpy(d["grid"]).cuda())
    info = {"img_idx": torch.zeros(256, 256, dtype=torch.long, device="cuda")}
    aff = m(rgb, info).reshape(256, 256, 3, 4)
    y = (aff[..., :3, :3] @ rgb[..., None] + aff[..., :3, 3:])[..., 0]  # scene_graph.py:95-98
    (y * G).sum().backward()
    st = int(d["stride"])
    assert np.abs(y.detach().cpu().numpy()[::st, ::st] - d["out"]).max() < 1e-5
    assert _rel(rgb.grad.cpu()[::st, ::st], torch.from_numpy(d["vrgb"])) < 1e-3
    assert _rel(m.bil_grids.grids.grad.cpu(), torch.from_numpy(d["vgrid"])) < 1e-3
    # fused transform
    rgb2 = rgb.detach().clone().requires_grad_(True)
    m.zero_grad()
    y2 = m.transform(rgb2, info)
    (y2 * G).sum().backward()
    assert (y2 - y).abs().max() < 1e-5
    assert _rel(rgb2.grad.cpu(), rgb.grad.cpu()) < 1e-3


@pytest.mark.parametrize("H,W", [(1080, 1920), (135, 241), (17, 16)])
@pytest.mark.parametrize("gf", [(4, 4, 2), None])
def test_against_oracle(H, W, gf):
    from bilateral_driving_b200 import synthetic as S
    from bilateral_driving_b200.bilateral import multiscale_bilateral
    from oracle import bilateral_ref as B

    sizes = S.GRID_SIZES_BASELINE
    grids = [g[1] for g in S.make_grids(2, sizes)]
    gen = torch.Generator(); gen.manual_seed(5)
    rgb = torch.rand(H, W, 3, generator=gen) * 1.2 - 0.1
    G = torch.randn(H, W, 3, generator=gen)
    big = H * W > 500_000
    # oracle (fp32 on CPU; fp64 for the small cases)
    dt = torch.float32 if big else torch.float64
    o_rgb = rgb.detach().clone().to(dt).requires_grad_(True)
    o_grids = [g.detach().clone().to(dt).requires_grad_(True) for g in grids]
    o_y = B.multiscale_forward(o_grids, o_rgb, gf)
    (o_y * G.to(dt)).sum().backward()
    c_rgb = rgb.detach().clone().cuda().requires_grad_(True)
    c_grids = [g.detach().clone().cuda().requires_grad_(True) for g in grids]
    c_y = multiscale_bilateral(c_rgb, c_grids, sizes, gf)
    (c_y * G.cuda()).sum().backward()
    # 1e-5 abs where the guidance resampling ratio is an integer (1080p: 270x480 / 540x960, the
    # reference's case); at non-integer ratios the fp32 tap weights of ANY implementation (torch's
    # included) carry ~1e-5 rounding, which this per-pixel-random guidance image amplifies
    exact_ratio = gf is None or all(H % f == 0 and W % f == 0 for f in gf)
    assert (c_y.detach().cpu() - o_y.detach().float()).abs().max() < (1e-5 if exact_ratio else 1e-4)
    # d/d(rgb) jumps where the guidance coordinate crosses a lattice plane: exclude those pixels
    amb = B.guidance_ambiguous(rgb, grids, gf)
    assert float(amb.float().mean()) < 0.02
    keep = (~amb)[..., None]
    assert _rel(c_rgb.grad.cpu() * keep, o_rgb.grad.float() * keep) < 1e-3
    for a, b in zip(c_grids, o_grids):
        assert _rel(a.grad.cpu(), b.grad.float()) < 1e-3


def test_module_api_and_state_dict():
    from bilateral_driving_b200.bilateral import MultiScaleBilateralAffineTransform

    m = MultiScaleBilateralAffineTransform("Affine", n=3, grid=[[2, 2, 1], [4, 4, 2], [8, 8, 4]]).cuda()
    assert sorted(m.state_dict().keys()) == sorted([
        "rgb2gray_weight", "bil_grids0.grids", "bil_grids0.rgb2gray_weight", "bil_grids1.grids",
        "bil_grids1.rgb2gray_weight", "bil_grids2.grids", "bil_grids2.rgb2gray_weight"])
    assert list(m.get_param_groups().keys()) == ["Affine#grid0", "Affine#grid1", "Affine#grid2"]
    rgb = torch.rand(40, 64, 3, device="cuda")
    info = {"img_idx": torch.full((40, 64), 2, dtype=torch.long, device="cuda")}
    out = m(rgb, info)
    assert [tuple(o.shape) for o in out] == [(1, 40, 64, 3, 4)] * 3
    # identity grids -> identity transform
    assert (m.transform(rgb, info) - rgb).abs().max() < 1e-6
    assert float(m.tv_loss()) == 0.0
    loss = m.inverse_loss(rgb, rgb)
    assert float(loss) < 1e-6
    # test-time branch
    m.in_test_set = True
    m.training_indices_for_test = {2: [0, 1]}
    assert (m.transform(rgb, info) - rgb).abs().max() < 1e-6


def test_tv_loss_matches_oracle():
    from bilateral_driving_b200.bilateral import total_variation_loss
    from oracle import bilateral_ref as B

    g = torch.randn(3, 12, 4, 5, 6)
    go = g.double().requires_grad_(True)
    B.total_variation_loss(go).backward()
    gc = g.cuda().requires_grad_(True)
    tv = total_variation_loss(gc)
    tv.backward()
    assert abs(float(tv) - float(B.total_variation_loss(go))) < 1e-5 * float(B.total_variation_loss(go))
    assert _rel(gc.grad.cpu(), go.grad.float()) < 1e-4


def test_generic_slice_arbitrary_xy():
    """BilateralGrid.forward / slice() with arbitrary xy samples (lib_bilagrid.py:171-230, 317-368)."""
    from bilateral_driving_b200.bilateral import BilateralGrid, slice as bil_slice
    from oracle import bilateral_ref as B

    torch.manual_seed(3)
    bg = BilateralGrid(num=3, grid_X=6, grid_Y=5, grid_W=4).cuda()
    with torch.no_grad():
        bg.grids.add_(0.1 * torch.randn_like(bg.grids))
    n = 500
    xy = torch.rand(n, 2) * 1.2 - 0.1          # some outside [0,1]: border padding
    rgb = torch.rand(n, 3) * 1.4 - 0.2
    idx = torch.full((n, 1), 2, dtype=torch.long)
    c_rgb = rgb.cuda().requires_grad_(True)
    out = bil_slice(bg, xy.cuda(), c_rgb, idx.cuda())
    G = torch.randn(n, 3, 4)
    (out["rgb_affine_mats"] * G.cuda()).sum().backward()
    # oracle
    g = bg.grids[2].detach().cpu().double().requires_grad_(True)
    o_rgb = rgb.double().requires_grad_(True)
    fx = xy[:, 0].double() * 5
    fy = xy[:, 1].double() * 4
    fz = B.luma_of(o_rgb) * 3
    ref = B.trilerp(g, fx, fy, fz).reshape(n, 3, 4)
    (ref * G.double()).sum().backward()
    assert (out["rgb_affine_mats"].detach().cpu() - ref.detach().float()).abs().max() < 1e-5
    ref_rgb = (ref[..., :3] @ o_rgb[..., None])[..., 0] + ref[..., 3]
    assert (out["rgb"].detach().cpu() - ref_rgb.detach().float()).abs().max() < 1e-5
    assert _rel(bg.grids.grad[2].cpu(), g.grad.float()) < 1e-3
    amb = ((fz.detach() - fz.detach().round()).abs() < 1e-4)
    keep = (~amb)[:, None].float()
    assert _rel(c_rgb.grad.cpu() * keep, o_rgb.grad.float() * keep) < 1e-3
