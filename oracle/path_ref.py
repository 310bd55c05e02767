"""CPU restatement of the WHOLE hot path of one reference training step for one rig of cameras:
activations + SH colour (vanilla.py:383-395) -> gsplat rasterization RGB+ED (base.py:393-408) ->
clamp(max=1) (base.py:417) -> sky composite (scene_graph.py:287-294) -> multi-scale bilateral slice
(modules.py:505-584) -> sequential 3x4 apply (scene_graph.py:112-117).  TEST INFRASTRUCTURE ONLY.
Rasteriser half PARITY UNPINNED (see oracle/raster_ref.py)."""
import torch

from . import bilateral_ref as B
from . import raster_ref as R
from . import sh_ref


def activate(params):
    q = params["_quats"]
    return dict(means=params["_means"], scales=torch.exp(params["_scales"]), quats=q / q.norm(dim=-1, keepdim=True),
                opacities=torch.sigmoid(params["_opacities"].reshape(-1)))


def sh_colors(params, viewmats, degree):
    coeffs = torch.cat([params["_features_dc"][:, None, :], params["_features_rest"]], dim=1)
    cols = []
    for c in range(viewmats.shape[0]):
        campos = torch.linalg.inv(viewmats[c])[:3, 3]
        dirs = params["_means"].detach() - campos
        cols.append(torch.clamp(sh_ref.spherical_harmonics(degree, dirs, coeffs) + 0.5, 0.0, 1.0))
    return torch.stack(cols)


def render_path(params, viewmats, Ks, width, height, sky=None, grid_slots=None, guidance_factor=None, sh_degree=3,
                near_plane=0.1, margin=1e-5, row_range=None):
    """Returns dict(rgb, rgb_gaussians, depth, opacity: [C,H,W,*], original_rgb, ambiguous [C,H,W], info)."""
    a = activate(params)
    colors = sh_colors(params, viewmats, sh_degree)
    renders, alphas, info = R.rasterization(a["means"], a["quats"], a["scales"], a["opacities"], colors, viewmats, Ks,
                                            width, height, near_plane=near_plane, render_mode="RGB+ED", margin=margin,
                                            row_range=row_range)
    rgb_g = torch.clamp(renders[..., :3], max=1.0)
    depth = renders[..., 3:4]
    rgb_in = rgb_g if sky is None else rgb_g + sky * (1.0 - alphas)
    if grid_slots is None:
        rgb = rgb_in
    else:
        rgb = torch.stack([B.multiscale_forward(grid_slots[c], rgb_in[c], guidance_factor)
                           for c in range(viewmats.shape[0])])
    amb = info["ambiguous"]
    if grid_slots is not None:
        amb = amb | torch.stack([B.guidance_ambiguous(rgb_in[c].detach(), grid_slots[c], guidance_factor)
                                 for c in range(viewmats.shape[0])])
    return dict(rgb=rgb, rgb_gaussians=rgb_g, depth=depth, opacity=alphas, original_rgb=rgb_in,
                ambiguous=amb, info=info)
