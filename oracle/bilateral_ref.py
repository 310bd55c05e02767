"""CPU restatement (plain torch index arithmetic, fp32/fp64, autograd-differentiable) of the
reference's multi-scale bilateral-grid colour correction.  TEST INFRASTRUCTURE ONLY.

Follows, in the reference tree (/root/reference/project):

* ``bilateral/lib_bilagrid.py:317-368``  BilateralGrid.forward: xy in [0,1] -> [-1,1], z = luma*2-1,
  5-D ``F.grid_sample(align_corners=True, padding_mode="border")`` over grids ``(N,12,L,GY,GX)``,
  channel c = 4*row + col of the 3x4 affine (``:362-363``), luma = (0.299, 0.587, 0.114) (``:287-288``).
* ``bilateral/lib_bilagrid.py:171-230``  slice(): plumbing around the above.
* ``bilateral/lib_bilagrid.py:152-168``  total_variation_loss.
* ``models/modules.py:494-504``  get_sample_grid: bilinear down-sample (align_corners=False) of the
  guidance RGB to (H//f, W//f) and a ``linspace(0,1)`` xy lattice.
* ``models/modules.py:409-420``  fill_matrix_res: bilinear up-sample of the 12-channel affine field.
* ``models/modules.py:505-584``  MultiScaleBilateralAffineTransform.forward (train branch and the
  test-time neighbour-averaging branch ``:523-547``).
* ``models/modules.py:275-351``  BilateralAffineTransform.forward (single grid, full-res guidance).
* ``models/trainers/scene_graph.py:112-117``  sequential 3x4 apply (no clamp).

None of this uses grid_sample / interpolate: it is an independent restatement, pinned against the
reference's own outputs by tests/test_oracle_bilateral.py (golden vectors from oracle/make_golden.py).
"""
import torch

LUMA = (0.299, 0.587, 0.114)


def lin_src(out_size: int, in_size: int, dtype=torch.float32):
    """Source taps of torch's bilinear resize with align_corners=False (no antialias).

    s = max((d + 0.5) * in/out - 0.5, 0); i0 = floor(s); i1 = min(i0 + 1, in - 1); t = s - i0.
    """
    d = torch.arange(out_size, dtype=dtype)
    scale = torch.tensor(in_size, dtype=dtype) / out_size
    s = (scale * (d + 0.5) - 0.5).clamp(min=0)
    i0 = s.floor().long().clamp(max=in_size - 1)
    i1 = (i0 + 1).clamp(max=in_size - 1)
    t = s - i0.to(dtype)
    return i0, i1, t


def resize_bilinear(img, out_h: int, out_w: int):
    """img [H, W, C] -> [out_h, out_w, C]; separable bilinear, align_corners=False."""
    H, W, _ = img.shape
    if (H, W) == (out_h, out_w):
        return img
    y0, y1, ty = lin_src(out_h, H, img.dtype)
    x0, x1, tx = lin_src(out_w, W, img.dtype)
    ty = ty[:, None, None]
    tx = tx[None, :, None]
    top = img[y0][:, x0] * (1 - tx) + img[y0][:, x1] * tx
    bot = img[y1][:, x0] * (1 - tx) + img[y1][:, x1] * tx
    return top * (1 - ty) + bot * ty


def linspace01(n: int, dtype=torch.float32):
    return torch.linspace(0.0, 1.0, n, dtype=dtype)


def trilerp(grid, fx, fy, fz):
    """grid [12, L, GY, GX]; fx/fy/fz broadcastable coords in voxel units -> [..., 12].

    Border padding: coordinates clamped to the lattice, the +1 corner clamped to the last cell
    (its weight is then zero)."""
    _, L, GY, GX = grid.shape
    fx = fx.clamp(0, GX - 1)
    fy = fy.clamp(0, GY - 1)
    fz = fz.clamp(0, L - 1)
    fx, fy, fz = torch.broadcast_tensors(fx, fy, fz)
    x0 = fx.floor().long().clamp(max=GX - 1)
    y0 = fy.floor().long().clamp(max=GY - 1)
    z0 = fz.floor().long().clamp(max=L - 1)
    x1 = (x0 + 1).clamp(max=GX - 1)
    y1 = (y0 + 1).clamp(max=GY - 1)
    z1 = (z0 + 1).clamp(max=L - 1)
    tx = (fx - x0.to(fx.dtype))[..., None]
    ty = (fy - y0.to(fy.dtype))[..., None]
    tz = (fz - z0.to(fz.dtype))[..., None]
    g = grid.permute(1, 2, 3, 0)  # [L, GY, GX, 12]

    def corner(zi, yi, xi):
        return g[zi, yi, xi]

    c00 = corner(z0, y0, x0) * (1 - tx) + corner(z0, y0, x1) * tx
    c01 = corner(z0, y1, x0) * (1 - tx) + corner(z0, y1, x1) * tx
    c10 = corner(z1, y0, x0) * (1 - tx) + corner(z1, y0, x1) * tx
    c11 = corner(z1, y1, x0) * (1 - tx) + corner(z1, y1, x1) * tx
    c0 = c00 * (1 - ty) + c01 * ty
    c1 = c10 * (1 - ty) + c11 * ty
    return c0 * (1 - tz) + c1 * tz


def luma_of(rgb):
    w = torch.tensor(LUMA, dtype=rgb.dtype)
    return (rgb * w).sum(-1)


def slice_lattice(grid, guide_rgb):
    """One grid [12,L,GY,GX] sliced on the linspace(0,1) lattice of guide_rgb [h,w,3] -> [h,w,12]."""
    _, L, GY, GX = grid.shape
    h, w, _ = guide_rgb.shape
    dt = guide_rgb.dtype
    fx = linspace01(w, dt)[None, :] * (GX - 1)
    fy = linspace01(h, dt)[:, None] * (GY - 1)
    fz = luma_of(guide_rgb) * (L - 1)
    return trilerp(grid, fx, fy, fz)


def multiscale_affines(grids, rgb, guidance_factor=(4, 4, 2)):
    """grids: list of [12,L,GY,GX] (already the image's slot, or a neighbour average);
    rgb [H,W,3].  Returns list of [H,W,3,4] (modules.py:505-584, train branch)."""
    H, W, _ = rgb.shape
    out = []
    for lvl, g in enumerate(grids):
        if guidance_factor is None:
            a = slice_lattice(g, rgb)
        else:
            f = guidance_factor[lvl]
            low = resize_bilinear(rgb, H // f, W // f)
            a = resize_bilinear(slice_lattice(g, low), H, W)
        out.append(a.reshape(H, W, 3, 4))
    return out


def apply_chain(rgb, affines):
    """scene_graph.py:112-117: x <- A[:, :3] x + A[:, 3], level after level, no clamp."""
    x = rgb
    for a in affines:
        x = (a[..., :3] @ x[..., None])[..., 0] + a[..., 3]
    return x


def multiscale_forward(grids, rgb, guidance_factor=(4, 4, 2)):
    return apply_chain(rgb, multiscale_affines(grids, rgb, guidance_factor))


def average_grids(grids_full, idx_list):
    """Test-time branch (modules.py:523-538): the mean over neighbour images' slices equals the
    slice of the mean grid (slicing is linear in grid values at fixed coordinates)."""
    return [torch.stack([g[i] for i in idx_list]).mean(0) for g in grids_full]


def total_variation_loss(x):
    """lib_bilagrid.py:152-168 for x [B, C, L, GY, GX]."""
    tv = 0
    for ax in range(2, x.dim()):
        n = x.shape[ax]
        a = x.narrow(ax, 1, n - 1)
        b = x.narrow(ax, 0, n - 1)
        count = max(a[0].numel(), 1)
        tv = tv + ((a - b) ** 2).sum() / count
    return tv / x.shape[0]


def tv_weights(grid_sizes):
    """modules.py:445: 0.5 * sqrt(X*Y*L) per level."""
    return [0.5 * (g[0] * g[1] * g[2]) ** 0.5 for g in grid_sizes]


def identity_grid(L, GY, GX, dtype=torch.float32):
    """lib_bilagrid.py:291-311."""
    g = torch.zeros(12, L, GY, GX, dtype=dtype)
    g[0] = 1
    g[5] = 1
    g[10] = 1
    return g


def guidance_ambiguous(rgb, grids, guidance_factor=(4, 4, 2), margin=1e-4):
    """Pixels whose guidance gradient may legitimately differ between two fp32 implementations:
    the guidance coordinate fz = luma*(L-1) of some level sits within ``margin`` of a lattice plane
    (the slab difference dA/dz jumps there) or of the clamp bounds 0 / L-1 (the gradient switches
    off there).  For low-resolution guidance the full-res taps feeding such a low-res pixel are
    flagged.  Outputs and grid gradients are continuous across these planes; only d/d(rgb) is not."""
    H, W, _ = rgb.shape
    amb = torch.zeros(H, W, dtype=torch.bool)
    with torch.no_grad():
        for lvl, g in enumerate(grids):
            L = g.shape[1]
            if L <= 1:
                continue
            f = 1 if guidance_factor is None else guidance_factor[lvl]
            low = rgb if f == 1 else resize_bilinear(rgb, H // f, W // f)
            fz = luma_of(low) * (L - 1)
            near = ((fz - fz.round()).abs() < margin) & (fz > -margin) & (fz < L - 1 + margin)
            if f == 1:
                amb |= near
            else:
                y0, y1, _ = lin_src(H // f, H, rgb.dtype)
                x0, x1, _ = lin_src(W // f, W, rgb.dtype)
                ys, xs = torch.nonzero(near, as_tuple=True)
                for yy in (y0, y1):
                    for xx in (x0, x1):
                        amb[yy[ys], xx[xs]] = True
    return amb
