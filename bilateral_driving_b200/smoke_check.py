"""One tiny forward + backward of the fused render + bilateral path on cuda:0, checked against the
CPU oracle (used by __graft_entry__.smoke(); the oracle is only the checker)."""
import torch


def run(verbose: bool = False) -> None:
    from bilateral_driving_b200 import synthetic as S
    from bilateral_driving_b200.render import render_fused
    from oracle.path_ref import render_path

    sizes = ((4, 4, 2), (8, 8, 4))
    p = S.make_gaussians(600, extent=8.0, scale_mean=0.15)
    p["_means"][:, 2] *= 0.4
    W, H, Cn = 64, 48, 1
    vm, Ks = S.make_rig(Cn, W, H)
    grids = S.make_grids(Cn, sizes)
    sky, _ = S.make_images(Cn, H, W)
    o_p = {k: v.double().requires_grad_(True) for k, v in p.items()}
    o_g = [g.double().requires_grad_(True) for g in grids]
    o = render_path(o_p, vm.double(), Ks.double(), W, H, sky=sky.double(), grid_slots=[[g[0] for g in o_g]],
                    guidance_factor=None)
    keep = (~o["ambiguous"])[..., None]
    (o["rgb"] * keep).sum().backward()
    c_p = {k: v.cuda().requires_grad_(True) for k, v in p.items()}
    c_g = [g.cuda().requires_grad_(True) for g in grids]
    out = render_fused(c_p, vm.cuda(), Ks.cuda(), W, H, sky=sky.cuda().view(H, W, 3), grid_slots=[[g[0] for g in c_g]],
                       bil_sizes=sizes, near_plane=0.1)
    (out["rgb"].view(Cn, H, W, 3) * keep.cuda()).sum().backward()
    err = float(((out["rgb"].view(Cn, H, W, 3).cpu() - o["rgb"].float()).abs() * keep).max())
    gerr = max(float((c_p[k].grad.cpu() - o_p[k].grad.float()).abs().max() / o_p[k].grad.abs().max().clamp(min=1e-12))
               for k in c_p)
    if verbose:
        print(f"smoke: rgb max abs err {err:.2e}, grad max rel err {gerr:.2e}, n_isect {out['info']['n_isect']}")
    assert err < 1e-5, err
    assert gerr < 1e-3, gerr
