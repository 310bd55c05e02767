"""Masked re-render: time of the reference's way (a second full rasterization with opacities * mask,
base.py:392-419) against rasterize_masked() over the cached sorted lists.  One 1920x1080 camera, synthetic scene."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bilateral_driving_b200 import synthetic as S
from bilateral_driving_b200.render import rasterization, rasterize_masked, spherical_harmonics


def main(n=2_000_000, iters=10):
    dev = "cuda"
    p = {k: v.to(dev) for k, v in S.make_gaussians(n).items()}
    vm, Ks = S.make_rig(1, 1920, 1080)
    vm, Ks = vm[:1].to(dev), Ks[:1].to(dev)
    quats = p["_quats"] / p["_quats"].norm(dim=-1, keepdim=True)
    scales, opac = torch.exp(p["_scales"]), torch.sigmoid(p["_opacities"])
    cam_pos = torch.linalg.inv(vm[0])[:3, 3]
    coeffs = torch.cat([p["_features_dc"][:, None, :], p["_features_rest"]], dim=1)
    cols = torch.clamp(spherical_harmonics(3, p["_means"] - cam_pos, coeffs) + 0.5, 0.0, 1.0)
    kw = dict(viewmats=vm, Ks=Ks, width=1920, height=1080, packed=False, absgrad=True, near_plane=0.1,
              render_mode="RGB+ED")
    mask = (torch.rand(n, device=dev) < 0.3)
    res = {}
    with torch.no_grad():
        _, _, info = rasterization(p["_means"], quats, scales, opac, cols, **kw)
        for name, fn in (("full_rerasterization", lambda: rasterization(p["_means"], quats, scales, opac * mask, cols, **kw)),
                         ("rasterize_masked", lambda: rasterize_masked(info, mask))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            res[name + "_ms"] = a.elapsed_time(b) / iters
        r1, a1 = rasterize_masked(info, mask)
        r2, a2, _ = rasterization(p["_means"], quats, scales, opac * mask, cols, **kw)
        res["bit_identical"] = bool(torch.equal(r1, r2) and torch.equal(a1, a2))
    res.update(n_gaussians=n, n_isect=info["n_isect"], kept_fraction=0.3)
    print(json.dumps(res))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000)
