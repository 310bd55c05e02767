"""``gsplat.rendering.rasterization`` signature: CPU tensors -> oracle/raster_ref.py."""
import torch

from oracle import raster_ref as R


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane=0.01,
                  far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None, packed=True, tile_size=16,
                  backgrounds=None, render_mode="RGB", sparse_grad=False, absgrad=False, rasterize_mode="classic",
                  **kw):
    if means.is_cuda:
        from bilateral_driving_b200.render import rasterization as product

        return product(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane=near_plane,
                       far_plane=far_plane, radius_clip=radius_clip, eps2d=eps2d, sh_degree=sh_degree, packed=packed,
                       tile_size=tile_size, backgrounds=backgrounds, render_mode=render_mode, sparse_grad=sparse_grad,
                       absgrad=absgrad, rasterize_mode=rasterize_mode, **kw)
    assert sh_degree is None and tile_size == 16 and not sparse_grad and not kw
    W = int(width.item()) if torch.is_tensor(width) else int(width)
    H = int(height.item()) if torch.is_tensor(height) else int(height)
    renders, alphas, info = R.rasterization(means, quats, scales, opacities, colors, viewmats, Ks, W, H,
                                            near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip,
                                            eps2d=eps2d, backgrounds=backgrounds, render_mode=render_mode,
                                            rasterize_mode=rasterize_mode, absgrad=absgrad)
    if absgrad:
        info["means2d"].absgrad = info["absgrad"]   # filled while the backward runs, like gsplat's
    info["radii"] = info["radii"].to(torch.int32)
    info["n_cameras"] = viewmats.shape[0]
    return renders, alphas, info
