"""Mint the committed golden fixtures under tests/golden/.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference for the bilateral fixtures):

    python -m oracle.make_golden

* ``bilateral_ms.npz``   - outputs and gradients of the UNMODIFIED reference
  ``MultiScaleBilateralAffineTransform`` (+ the scene_graph.py:112-117 apply loop) on a small
  non-multiple-of-4 image, for ``guidance_factor=[4,4,2]`` and ``None``, train and test branch.
* ``bilateral_cfg1.npz`` - BASELINE.json config 1 (single 16x16x8 grid, 256x256 image) through the
  reference ``BilateralAffineTransform`` + apply of scene_graph.py:95-98; inputs are regenerated from
  seeds, outputs stored sub-sampled.
* ``raster_small.npz``   - outputs/gradients of oracle/raster_ref.py + sh_ref.py in fp64 on a small
  scene.  These pin the restatement against itself across refactors only: the rasteriser half is
  PARITY UNPINNED against gsplat (see oracle/__init__.py).
"""
import os

import numpy as np
import torch

from bilateral_driving_b200 import synthetic as S
from oracle import bilateral_ref as B
from oracle import raster_ref as R
from oracle import sh_ref
from oracle.ref_loader import load_reference, reference_apply_chain

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def bilateral_ms():
    _, mods = load_reference()
    sizes = [[4, 4, 2], [8, 8, 4], [6, 5, 3]]
    n, idx, H, W = 3, 1, 46, 70
    m = mods.MultiScaleBilateralAffineTransform("Affine", n=n, grid=sizes, device="cpu")
    for i in range(3):
        g = getattr(m, f"bil_grids{i}").grids
        g.data += 0.05 * torch.randn(g.shape, generator=gen(16 + i))
    rgb0 = torch.rand(H, W, 3, generator=gen(0)) * 1.3 - 0.1  # some luma outside [0,1]
    G = torch.randn(H, W, 3, generator=gen(2))
    info = {"img_idx": torch.full((H, W), idx, dtype=torch.long)}
    out = dict(sizes=np.array(sizes), n=n, idx=idx, rgb=rgb0.numpy(), G=G.numpy())
    for i in range(3):
        out[f"grids{i}"] = getattr(m, f"bil_grids{i}").grids.detach().numpy().copy()
    for tag, gf in (("f442", [4, 4, 2]), ("none", None)):
        rgb = rgb0.clone().requires_grad_(True)
        affs = m(rgb, info, guidance_factor=gf)
        y = reference_apply_chain(rgb, affs)
        (y * G).sum().backward()
        out[f"{tag}_out"] = y.detach().numpy()
        out[f"{tag}_aff2"] = affs[2].detach().numpy()[0].reshape(H, W, 12)
        out[f"{tag}_vrgb"] = rgb.grad.numpy().copy()
        for i in range(3):
            p = getattr(m, f"bil_grids{i}").grids
            out[f"{tag}_vgrid{i}"] = p.grad.numpy().copy()
            p.grad = None
    # test-time branch: neighbour averaging (modules.py:523-547)
    m.in_test_set = True
    m.training_indices_for_test = {idx: [0, 2]}
    with torch.no_grad():
        y = reference_apply_chain(rgb0, m(rgb0, info, guidance_factor=[4, 4, 2]))
    out["test_f442_out"] = y.numpy()
    out["tv"] = float(m.tv_loss())
    np.savez_compressed(os.path.join(OUT, "bilateral_ms.npz"), **out)
    print("bilateral_ms.npz written")


def bilateral_cfg1():
    _, mods = load_reference()
    H = W = 256
    m = mods.BilateralAffineTransform("Affine", n=1, grid_X=16, grid_Y=16, grid_W=8, device="cpu")
    g = m.bil_grids.grids
    g.data += 0.05 * torch.randn(g.shape, generator=gen(1))
    rgb = torch.rand(H, W, 3, generator=gen(0)).requires_grad_(True)
    G = torch.randn(H, W, 3, generator=gen(2))
    info = {"img_idx": torch.zeros(H, W, dtype=torch.long)}
    aff = m(rgb, info).reshape(H, W, 3, 4)
    y = (aff[..., :3, :3] @ rgb[..., None] + aff[..., :3, 3:])[..., 0]  # scene_graph.py:95-98
    (y * G).sum().backward()
    st = 5
    np.savez_compressed(
        os.path.join(OUT, "bilateral_cfg1.npz"), stride=st, out=y.detach().numpy()[::st, ::st],
        vrgb=rgb.grad.numpy()[::st, ::st], vgrid=g.grad.numpy().copy(), grid=g.detach().numpy().copy(),
        out_sum=float(y.sum()), vrgb_abs_sum=float(rgb.grad.abs().sum()))
    print("bilateral_cfg1.npz written")


def small_scene(dtype=torch.float64):
    """Shared by the fixture and by tests (tests regenerate inputs from the same seeds)."""
    p = S.make_gaussians(1500, extent=10.0, scale_mean=0.12)
    p["_means"][:, 2] = p["_means"][:, 2] * 0.5
    W, H = 88, 56
    vm, Ks = S.make_rig(2, W, H)
    return {k: v.to(dtype) for k, v in p.items()}, vm.to(dtype), Ks.to(dtype), W, H


def raster_small():
    p, vm, Ks, W, H = small_scene()
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    scales = torch.exp(leaves["_scales"])
    quats = leaves["_quats"]
    opac = torch.sigmoid(leaves["_opacities"])
    coeffs = torch.cat([leaves["_features_dc"][:, None], leaves["_features_rest"]], 1)
    cols = []
    for c in range(vm.shape[0]):
        campos = torch.linalg.inv(vm[c])[:3, 3]
        dirs = leaves["_means"].detach() - campos
        cols.append((sh_ref.spherical_harmonics(3, dirs, coeffs) + 0.5).clamp(0, 1))
    r, a, info = R.rasterization(leaves["_means"], quats, scales, opac, torch.stack(cols), vm, Ks, W, H,
                                 near_plane=0.1, render_mode="RGB+ED")
    keep = (~info["ambiguous"])[..., None].to(r.dtype)  # ambiguous pixels carry no loss
    Gr = torch.randn(r.shape, generator=gen(3), dtype=r.dtype) * keep
    Ga = torch.randn(a.shape, generator=gen(4), dtype=r.dtype) * keep
    ((r * Gr).sum() + (a * Ga).sum()).backward()
    out = dict(render=r.detach().numpy(), alpha=a.detach().numpy(), ambiguous=info["ambiguous"].numpy(),
               Gr=Gr.numpy(), Ga=Ga.numpy(),
               radii=info["radii"].numpy(), means2d=info["means2d"].detach().numpy(),
               conics=info["conics"].detach().numpy(), depths=info["depths"].detach().numpy(),
               n_isect=np.array(info["n_isect"]))
    for k, v in leaves.items():
        out["v" + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "raster_small.npz"), **out)
    print("raster_small.npz written; ambiguous px:", int(info["ambiguous"].sum()), "n_isect", info["n_isect"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    bilateral_ms()
    bilateral_cfg1()
    raster_small()
