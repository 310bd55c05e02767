"""Synthetic Gaussians / camera rig / grids of BASELINE.json's configs.

Everything is generated on the CPU with fixed seeds (so the distribution does not depend on the
device) and moved afterwards.  Parameter names mirror ``VanillaGaussians`` in the reference
(``models/gaussians/vanilla.py:60-110``): ``_means, _scales (log), _quats (raw), _opacities (logit),
_features_dc [N,3], _features_rest [N,15,3]``.
"""
import math
from typing import Dict, List, Sequence

import torch

SH_C0 = 0.28209479177387814
GRID_SIZES_BASELINE = ((8, 8, 4), (16, 16, 8), (32, 32, 16))       # BASELINE.json "8/16/32"
GRID_SIZES_REFERENCE = ((2, 2, 1), (4, 4, 2), (8, 8, 4))            # omnire_ms_bilateral.yaml:249


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_gaussians(n: int, extent: float = 60.0, scale_mean: float = 0.08, sh_rest_std: float = 0.05,
                   seed_offset: int = 0) -> Dict[str, torch.Tensor]:
    """Raw (pre-activation) parameters, fp32, CPU."""
    o = seed_offset
    xy = (torch.rand(n, 2, generator=_gen(10 + o)) * 2 - 1) * extent
    z = torch.rand(n, 1, generator=_gen(110 + o)) * 9.0 - 1.0
    means = torch.cat([xy, z], dim=1)
    scales = math.log(scale_mean) + 0.6 * torch.randn(n, 3, generator=_gen(11 + o))
    quats = torch.randn(n, 4, generator=_gen(12 + o))
    opac = 2.0 * torch.randn(n, generator=_gen(13 + o))
    fdc = (torch.rand(n, 3, generator=_gen(14 + o)) - 0.5) / SH_C0
    frest = sh_rest_std * torch.randn(n, 15, 3, generator=_gen(15 + o))
    return dict(_means=means, _scales=scales, _quats=quats, _opacities=opac,
                _features_dc=fdc, _features_rest=frest)


def look_at_yaw(yaw_deg: float, origin=(0.0, 0.0, 1.5)) -> torch.Tensor:
    """cam-to-world of an OpenCV camera (x right, y down, z forward) in a z-up world, looking
    along the horizontal direction rotated ``yaw_deg`` from +x.  Returns [4,4] fp32."""
    yaw = math.radians(yaw_deg)
    fwd = torch.tensor([math.cos(yaw), math.sin(yaw), 0.0])
    down = torch.tensor([0.0, 0.0, -1.0])
    right = torch.linalg.cross(down, fwd)  # x = y cross z
    c2w = torch.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2] = right, down, fwd
    c2w[:3, 3] = torch.tensor(origin)
    return c2w


def make_rig(n_cams: int = 6, width: int = 1920, height: int = 1080, focal_scale: float = 1.0):
    """nuScenes-shaped rig (configs/datasets/nuscenes/6cams.yaml:3-9).  Returns viewmats [C,4,4]
    (world-to-camera), Ks [C,3,3]."""
    yaws = [0.0, 55.0, -55.0, 110.0, -110.0, 180.0][:n_cams]
    offs = [(1.5, 0.0), (1.5, 0.5), (1.5, -0.5), (0.0, 0.5), (0.0, -0.5), (-0.5, 0.0)][:n_cams]
    viewmats, Ks = [], []
    for i, yaw in enumerate(yaws):
        f = (970.0 if i == 5 else 1520.0) * focal_scale * (width / 1920.0)
        c2w = look_at_yaw(yaw, (offs[i][0], offs[i][1], 1.5))
        viewmats.append(torch.linalg.inv(c2w))
        Ks.append(torch.tensor([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]]))
    return torch.stack(viewmats), torch.stack(Ks)


def make_grids(n_images: int, sizes: Sequence[Sequence[int]] = GRID_SIZES_BASELINE,
               noise: float = 0.05) -> List[torch.Tensor]:
    """Identity + noise grids in the reference layout (N,12,L,GY,GX) (lib_bilagrid.py:283-311)."""
    out = []
    for lvl, (gx, gy, gl) in enumerate(sizes):
        g = torch.zeros(n_images, 12, gl, gy, gx)
        g[:, 0] = 1.0
        g[:, 5] = 1.0
        g[:, 10] = 1.0
        g = g + noise * torch.randn(g.shape, generator=_gen(16 + 100 * lvl))
        out.append(g)
    return out


def make_images(n_cams: int, height: int, width: int):
    sky = torch.rand(n_cams, height, width, 3, generator=_gen(17))
    gt = torch.rand(n_cams, height, width, 3, generator=_gen(18))
    return sky, gt


def activate(params: Dict[str, torch.Tensor]):
    """vanilla.py:122-146 activations (reference semantics, plain torch)."""
    return dict(
        means=params["_means"],
        scales=torch.exp(params["_scales"]),
        quats=params["_quats"] / params["_quats"].norm(dim=-1, keepdim=True),
        opacities=torch.sigmoid(params["_opacities"]),
    )
