"""CPU: the bilateral restatement (oracle/bilateral_ref.py) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py), and - when /root/reference is present - against the
reference itself in fp64."""
import os

import numpy as np
import pytest
import torch

from oracle import bilateral_ref as B
from oracle.ref_loader import reference_available


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _slots(d, idx):
    return [torch.from_numpy(d[f"grids{i}"])[idx].clone().requires_grad_(True) for i in range(3)]


@pytest.mark.parametrize("tag,gf", [("f442", (4, 4, 2)), ("none", None)])
def test_multiscale_matches_reference_golden(golden_dir, tag, gf):
    d = _load(golden_dir, "bilateral_ms.npz")
    idx = int(d["idx"])
    rgb = torch.from_numpy(d["rgb"]).clone().requires_grad_(True)
    G = torch.from_numpy(d["G"])
    slots = _slots(d, idx)
    affs = B.multiscale_affines(slots, rgb, gf)
    y = B.apply_chain(rgb, affs)
    (y * G).sum().backward()
    assert np.abs(y.detach().numpy() - d[f"{tag}_out"]).max() < 2e-6
    assert np.abs(affs[2].detach().numpy().reshape(*rgb.shape[:2], 12) - d[f"{tag}_aff2"]).max() < 2e-6
    assert np.abs(rgb.grad.numpy() - d[f"{tag}_vrgb"]).max() < 5e-5
    for i in range(3):
        ref = d[f"{tag}_vgrid{i}"]
        assert np.abs(slots[i].grad.numpy() - ref[idx]).max() < 5e-5 * max(1.0, np.abs(ref).max())
        other = np.delete(ref, idx, axis=0)
        assert np.abs(other).max() == 0.0  # advanced indexing -> dense grad, zero outside the slot


def test_test_time_neighbour_average(golden_dir):
    d = _load(golden_dir, "bilateral_ms.npz")
    full = [torch.from_numpy(d[f"grids{i}"]) for i in range(3)]
    y = B.multiscale_forward(B.average_grids(full, [0, 2]), torch.from_numpy(d["rgb"]), (4, 4, 2))
    assert np.abs(y.numpy() - d["test_f442_out"]).max() < 2e-6


def test_tv_loss(golden_dir):
    d = _load(golden_dir, "bilateral_ms.npz")
    w = B.tv_weights(d["sizes"].tolist())
    tv = sum(wi * B.total_variation_loss(torch.from_numpy(d[f"grids{i}"])) for i, wi in enumerate(w))
    assert abs(float(tv) - float(d["tv"])) < 1e-5 * abs(float(d["tv"]))


def test_config1_single_grid(golden_dir):
    """BASELINE.json configs[0]: single 16x16x8 grid slice+apply on a 256x256 image (CPU plumbing)."""
    d = _load(golden_dir, "bilateral_cfg1.npz")
    g0 = torch.Generator(); g0.manual_seed(0)
    g2 = torch.Generator(); g2.manual_seed(2)
    rgb = torch.rand(256, 256, 3, generator=g0).requires_grad_(True)
    G = torch.randn(256, 256, 3, generator=g2)
    slot = torch.from_numpy(d["grid"])[0].clone().requires_grad_(True)
    y = B.multiscale_forward([slot], rgb, None)
    (y * G).sum().backward()
    st = int(d["stride"])
    assert np.abs(y.detach().numpy()[::st, ::st] - d["out"]).max() < 2e-6
    assert np.abs(rgb.grad.numpy()[::st, ::st] - d["vrgb"]).max() < 5e-5
    assert np.abs(slot.grad.numpy() - d["vgrid"][0]).max() < 2e-4 * np.abs(d["vgrid"]).max()
    assert abs(float(y.detach().sum()) - float(d["out_sum"])) < 1e-4 * abs(float(d["out_sum"]))


def test_edge_cases():
    # L == 1 grid: no guidance gradient; luma far outside [0,1]: clamped, zero guidance gradient
    g = (B.identity_grid(1, 3, 2) + 0.1 * torch.randn(12, 1, 3, 2)).requires_grad_(True)
    rgb = (torch.rand(9, 7, 3) * 4 - 2).requires_grad_(True)
    y = B.multiscale_forward([g], rgb, None)
    y.sum().backward()
    assert torch.isfinite(rgb.grad).all() and torch.isfinite(g.grad).all()
    # 1x1 low-res lattice (H//f == 1)
    y = B.multiscale_forward([B.identity_grid(2, 2, 2)], torch.rand(5, 6, 3), (4,))
    assert y.shape == (5, 6, 3)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_fp64_against_live_reference():
    from oracle.ref_loader import load_reference, reference_apply_chain

    _, mods = load_reference()
    torch.manual_seed(0)
    sizes = [[8, 8, 4], [16, 16, 8], [32, 32, 16]]
    torch.set_default_dtype(torch.float64)  # the module creates its luma buffer / lattice in the default dtype
    try:
        m = mods.MultiScaleBilateralAffineTransform("Affine", n=2, grid=sizes, device="cpu").double()
        for i in range(3):
            p = getattr(m, f"bil_grids{i}").grids
            p.data += 0.05 * torch.randn_like(p)
        rgb = torch.rand(37, 53, 3) * 1.2 - 0.1
        info = {"img_idx": torch.ones(37, 53, dtype=torch.long)}
        for gf in ([4, 4, 2], None):
            ref = reference_apply_chain(rgb, m(rgb, info, guidance_factor=gf))
            ours = B.multiscale_forward([getattr(m, f"bil_grids{i}").grids[1].detach() for i in range(3)], rgb, gf)
            assert (ref - ours).abs().max() < 1e-12
    finally:
        torch.set_default_dtype(torch.float32)


def test_module_mirror_keeps_the_reference_checkpoint_contract():
    """State-dict keys / shapes / dtypes and optimiser group names of the host mirror equal the UNMODIFIED
    reference modules (models/modules.py:275-351, 422-593): released checkpoints load with strict=True
    (models/trainers/base.py:737) and the YAML optimiser entries `Affine#grid{i}` keep matching (base.py:182-188)."""
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    from bilateral_driving_b200 import bilateral as ours

    _, mods = load_reference()
    sizes = [[2, 2, 1], [4, 4, 2], [8, 8, 4]]   # configs/omnire_ms_bilateral.yaml:249
    ref = mods.MultiScaleBilateralAffineTransform("Affine", n=5, grid=sizes, device="cpu")
    mine = ours.MultiScaleBilateralAffineTransform("Affine", n=5, grid=sizes, device="cpu")
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys())
    for k in sd_ref:
        assert sd_ref[k].shape == sd_mine[k].shape and sd_ref[k].dtype == sd_mine[k].dtype, k
        assert torch.equal(sd_ref[k], sd_mine[k]), k          # identity initialisation, luma weights
    mine.load_state_dict(sd_ref, strict=True)
    g_ref, g_mine = ref.get_param_groups(), mine.get_param_groups()
    assert list(g_ref.keys()) == list(g_mine.keys())
    for k in g_ref:
        assert [tuple(p.shape) for p in g_ref[k]] == [tuple(p.shape) for p in g_mine[k]]
    assert [float(w) for w in ref.tv_weight] == [float(w) for w in mine.tv_weight]
    # single-scale module (models/modules.py:275-351)
    r1 = mods.BilateralAffineTransform("Affine", n=3, grid_X=16, grid_Y=16, grid_W=8, device="cpu")
    m1 = ours.BilateralAffineTransform("Affine", n=3, grid_X=16, grid_Y=16, grid_W=8, device="cpu")
    assert list(r1.state_dict().keys()) == list(m1.state_dict().keys())
    for k, v in r1.state_dict().items():
        assert v.shape == m1.state_dict()[k].shape, k
    assert list(r1.get_param_groups().keys()) == list(m1.get_param_groups().keys())
