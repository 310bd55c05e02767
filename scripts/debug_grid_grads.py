"""Debug: fused mode-2 grid gradients vs the stand-alone bilateral backward on the same image."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bilateral_driving_b200 import synthetic as S
from bilateral_driving_b200.render import render_fused
from bilateral_driving_b200.bilateral import multiscale_bilateral
from oracle.make_golden import small_scene

SIZES = ((4, 4, 2), (8, 8, 4), (6, 5, 3))
p, vm, Ks, W, H = small_scene(torch.float32)
Cn = 2
grids = S.make_grids(Cn, SIZES)
sky, _ = S.make_images(Cn, H, W)
c_p = {k: v.cuda() for k, v in p.items()}
c_g = [g.cuda().requires_grad_(True) for g in grids]
slots = [[g[c] for g in c_g] for c in range(Cn)]
out = render_fused(c_p, vm.cuda(), Ks.cuda(), W, H, sky=sky.cuda().view(Cn * H, W, 3), grid_slots=slots, bil_sizes=SIZES, near_plane=0.1)
gen = torch.Generator(); gen.manual_seed(7)
G = torch.randn(Cn * H, W, 3, generator=gen).cuda()
(out["rgb"] * G).sum().backward()
fused = [g.grad.clone() for g in c_g]
# stand-alone on the same pre-affine image
rgb_in = (out["rgb_gaussians"] + sky.cuda().view(Cn * H, W, 3) * (1 - out["opacity"])).detach().view(Cn, H, W, 3)
g2 = [g.detach().clone().requires_grad_(True) for g in c_g]
for c in range(Cn):
    y = multiscale_bilateral(rgb_in[c], [g[c] for g in g2], SIZES, None)
    (y * G.view(Cn, H, W, 3)[c]).sum().backward()
for l in range(3):
    d = (fused[l] - g2[l].grad).abs()
    print("level", l, SIZES[l], "max abs diff", float(d.max()), "max ref", float(g2[l].grad.abs().max()))
    bad = torch.nonzero(d > 1e-3 * g2[l].grad.abs().max())
    print("  n bad", bad.shape[0], "first", bad[:12].tolist())
    if bad.shape[0]:
        nodes = torch.unique(bad[:, [0, 2, 3, 4]], dim=0)
        print("  bad nodes (cam,z,y,x):", nodes[:20].tolist())
        i = tuple(bad[0].tolist())
        print("  e.g.", i, float(fused[l][i]), float(g2[l].grad[i]))
