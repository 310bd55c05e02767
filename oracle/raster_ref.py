"""CPU restatement (plain torch, fp32/fp64, autograd-differentiable) of the gsplat v1.3.0
``rasterization`` semantics the reference calls at ``models/trainers/base.py:393-408``.
TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic lives in gsplat v1.3.0 (pip dependency, README.md:81 of the
reference; import seam models/gaussians/basics.py:12-15).  It is not vendored under
/root/reference, not installed and not installable here (no network), and the reference holds no
tests / golden vectors at this boundary.  This file restates the published algorithm as recorded in
SURVEY.md section 8c: projection (EWA, eps2d blur 0.3, symmetric 1.3*tan(fov) clamp,
radius = ceil(3 sqrt(lambda_max))), 16x16 tile binning on the square bound, (tile, fp32 depth bits)
ordering with ties broken by Gaussian id, front-to-back blend with alpha < 1/255 skip,
alpha <= 0.999 clamp, stop BEFORE the Gaussian that would push T <= 1e-4, and the RGB+ED depth
normalisation.

Everything is vectorised per tile so gradients come from torch autograd (the masks are constants,
exactly as in the CUDA backward of gsplat).  Besides outputs it returns an ``ambiguous`` pixel mask:
pixels where some threshold decision (alpha vs 1/255, T' vs 1e-4, ceil/floor in the radius / tile
bound) is within a relative margin of flipping, so an fp32 implementation may legitimately differ
there.  Parity tests compare all other pixels at the stated tolerance.
"""
import math

import torch

TILE = 16
ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.999
T_STOP = 1e-4


def quat_to_rotmat(quats):
    """wxyz quaternion (normalised inside) -> [...,3,3]."""
    q = quats / quats.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        dim=-1,
    )
    return R.reshape(quats.shape[:-1] + (3, 3))


def project(means, quats, scales, viewmat, K, width, height, eps2d=0.3, near_plane=0.01,
            far_plane=1e10, radius_clip=0.0, margin=1e-5):
    """One camera.  Returns dict(radii[N] int64, means2d[N,2], depths[N], conics[N,3],
    compensations[N], ambiguous[N] bool)."""
    dt = means.dtype
    R = viewmat[:3, :3]
    t = viewmat[:3, 3]
    pc = means @ R.T + t  # [N,3]
    x, y, z = pc.unbind(-1)
    valid = (z >= near_plane) & (z <= far_plane)
    zs = torch.where(valid, z, torch.ones_like(z))  # keep culled rows finite

    Rq = quat_to_rotmat(quats)
    M = Rq * scales[:, None, :]
    cov = M @ M.transpose(-1, -2)
    covc = R @ cov @ R.T

    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    limx = 1.3 * (0.5 * width / fx)
    limy = 1.3 * (0.5 * height / fy)
    rz = 1.0 / zs
    tx = zs * torch.minimum(limx, torch.maximum(-limx, x * rz))
    ty = zs * torch.minimum(limy, torch.maximum(-limy, y * rz))
    zero = torch.zeros_like(zs)
    J = torch.stack(
        [fx * rz, zero, -fx * tx * rz * rz, zero, fy * rz, -fy * ty * rz * rz], dim=-1
    ).reshape(-1, 2, 3)
    cov2 = J @ covc @ J.transpose(-1, -2)
    means2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], dim=-1)

    a0, b0, c0 = cov2[:, 0, 0], cov2[:, 0, 1], cov2[:, 1, 1]
    det0 = a0 * c0 - b0 * b0
    a, c = a0 + eps2d, c0 + eps2d
    det = a * c - b0 * b0
    valid = valid & (det > 0)
    dets = torch.where(valid, det, torch.ones_like(det))
    conics = torch.stack([c / dets, -b0 / dets, a / dets], dim=-1)
    comp = torch.sqrt(torch.clamp(det0 / dets, min=0))

    with torch.no_grad():
        bh = 0.5 * (a + c)
        lam = bh + torch.sqrt(torch.clamp(bh * bh - det, min=0.01))
        r_f = 3.0 * torch.sqrt(lam)
        radius = torch.ceil(r_f)
        amb = (r_f - torch.floor(r_f)).abs() < margin * r_f.clamp(min=1)
        amb |= (torch.ceil(r_f) - r_f).abs() < margin * r_f.clamp(min=1)
        valid = valid & (radius > radius_clip)
        mx, my = means2d[:, 0], means2d[:, 1]
        valid = valid & ~((mx + radius <= 0) | (mx - radius >= width) | (my + radius <= 0) | (my - radius >= height))
        radii = torch.where(valid, radius, torch.zeros_like(radius)).long()
        # tile-bound rounding ambiguity
        for v in (mx / TILE - radius / TILE, mx / TILE + radius / TILE, my / TILE - radius / TILE, my / TILE + radius / TILE):
            amb |= (v - torch.round(v)).abs() < margin * v.abs().clamp(min=1)
        amb &= valid
    return dict(radii=radii, means2d=means2d, depths=z, conics=conics, compensations=comp.to(dt), ambiguous=amb)


def tile_rect(means2d, radii, width, height):
    """isect_tiles bound: tile_min inclusive, tile_max exclusive, clamped to the tile lattice."""
    tw = (width + TILE - 1) // TILE
    th = (height + TILE - 1) // TILE
    r = radii.to(means2d.dtype) / TILE
    cxy = means2d.detach() / TILE
    tmin_x = torch.floor(cxy[:, 0] - r).clamp(0, tw).long()
    tmin_y = torch.floor(cxy[:, 1] - r).clamp(0, th).long()
    tmax_x = torch.ceil(cxy[:, 0] + r).clamp(0, tw).long()
    tmax_y = torch.ceil(cxy[:, 1] + r).clamp(0, th).long()
    vis = radii > 0
    z = torch.zeros_like(tmin_x)
    return (torch.where(vis, tmin_x, z), torch.where(vis, tmin_y, z),
            torch.where(vis, tmax_x, z), torch.where(vis, tmax_y, z))


def depth_order(depths):
    """Stable order by the fp32 bit pattern of depth (positive floats order like their bits);
    ties by Gaussian id, as the stable radix sort over emission order gives."""
    key = depths.detach().to(torch.float32)
    return torch.sort(key, stable=True).indices


def composite(means2d, conics, opacities, colors, depths, radii, width, height, background=None,
              margin=1e-5, row_range=None, abs_acc=None):
    """One camera.  colors [N,D].  Returns render [H,W,D], alpha [H,W], ambiguous [H,W] bool,
    n_isect (int).  ``row_range`` = (tile_row_begin, tile_row_end) restricts the tiles rendered.
    ``abs_acc`` [N,2]: filled DURING backward with gsplat's ``absgrad`` = sum over pixels of |per-pixel
    contribution to v_means2d|, which plain autograd (a signed sum) cannot give: the
    per-(pixel, Gaussian) cotangent of the expanded means is caught by a tensor hook before it is reduced."""
    dt = means2d.dtype
    D = colors.shape[-1]
    tw = (width + TILE - 1) // TILE
    th = (height + TILE - 1) // TILE
    tmin_x, tmin_y, tmax_x, tmax_y = tile_rect(means2d, radii, width, height)
    order = depth_order(depths)
    render = torch.zeros(height, width, D, dtype=dt)
    alpha_img = torch.zeros(height, width, dtype=dt)
    amb_img = torch.zeros(height, width, dtype=torch.bool)
    n_isect = int(((tmax_x - tmin_x) * (tmax_y - tmin_y)).sum())
    rows = range(th) if row_range is None else range(row_range[0], row_range[1])
    out_rows = []
    for tyi in rows:
        row_r, row_a, row_m = [], [], []
        y0, y1 = tyi * TILE, min((tyi + 1) * TILE, height)
        for txi in range(tw):
            x0, x1 = txi * TILE, min((txi + 1) * TILE, width)
            inside = (tmin_x <= txi) & (txi < tmax_x) & (tmin_y <= tyi) & (tyi < tmax_y)
            ids = order[inside[order]]
            py, px = torch.meshgrid(
                torch.arange(y0, y1, dtype=dt) + 0.5, torch.arange(x0, x1, dtype=dt) + 0.5, indexing="ij")
            P = py.numel()
            if ids.numel() == 0:
                r = torch.zeros(P, D, dtype=dt)
                a = torch.zeros(P, dtype=dt)
                m = torch.zeros(P, dtype=torch.bool)
                Tf = torch.ones(P, dtype=dt)
            else:
                mu = means2d[ids]
                con = conics[ids]
                op = opacities[ids]
                mu_e = mu[None, :, :].expand(P, -1, -1)
                if abs_acc is not None and mu_e.requires_grad:
                    def _catch(g, ids=ids):
                        abs_acc.index_add_(0, ids, g.detach().abs().sum(0))

                    mu_e.register_hook(_catch)
                dx = mu_e[:, :, 0] - px.reshape(-1, 1)
                dy = mu_e[:, :, 1] - py.reshape(-1, 1)
                sigma = 0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) + con[None, :, 1] * dx * dy
                araw = op[None, :] * torch.exp(-sigma)
                alpha = torch.clamp(araw, max=ALPHA_MAX)
                with torch.no_grad():
                    skip = (sigma < 0) | (alpha < ALPHA_MIN)
                    a_eff = torch.where(skip, torch.zeros_like(alpha), alpha)
                    T_after = torch.cumprod(1 - a_eff, dim=1)
                    stopped = torch.cummax((T_after <= T_STOP).to(torch.int8), dim=1).values.bool()
                    include = ~skip & ~stopped
                    # ambiguity: decisions evaluated before the stop that sit on a threshold
                    live = ~torch.cat([torch.zeros(P, 1, dtype=torch.bool), stopped[:, :-1]], dim=1)
                    near_a = ((alpha - ALPHA_MIN).abs() < margin * ALPHA_MIN * 10) & live
                    near_t = ((T_after - T_STOP).abs() < margin * T_STOP * 10) & live & ~skip
                    near_c = ((araw - ALPHA_MAX).abs() < margin) & live  # clamp kink: gradient only
                    m = (near_a | near_t | near_c).any(dim=1)
                a_inc = alpha * include.to(dt)
                T_before = torch.cumprod(torch.cat([torch.ones(P, 1, dtype=dt), (1 - a_inc)[:, :-1]], dim=1), dim=1)
                w = a_inc * T_before
                r = w @ colors[ids]
                Tf = torch.prod(1 - a_inc, dim=1)
                a = 1 - Tf
            if background is not None:
                r = r + Tf[:, None] * background[None, :]
            row_r.append(r.reshape(y1 - y0, x1 - x0, D))
            row_a.append(a.reshape(y1 - y0, x1 - x0))
            row_m.append(m.reshape(y1 - y0, x1 - x0))
        out_rows.append((y0, y1, torch.cat(row_r, 1), torch.cat(row_a, 1), torch.cat(row_m, 1)))
    if out_rows:
        # assemble without in-place writes so autograd stays simple
        pieces_r, pieces_a, pieces_m = [], [], []
        cur = 0
        for (y0, y1, r, a, m) in out_rows:
            if y0 > cur:
                pieces_r.append(torch.zeros(y0 - cur, width, D, dtype=dt))
                pieces_a.append(torch.zeros(y0 - cur, width, dtype=dt))
                pieces_m.append(torch.zeros(y0 - cur, width, dtype=torch.bool))
            pieces_r.append(r); pieces_a.append(a); pieces_m.append(m)
            cur = y1
        if cur < height:
            pieces_r.append(torch.zeros(height - cur, width, D, dtype=dt))
            pieces_a.append(torch.zeros(height - cur, width, dtype=dt))
            pieces_m.append(torch.zeros(height - cur, width, dtype=torch.bool))
        render, alpha_img, amb_img = torch.cat(pieces_r, 0), torch.cat(pieces_a, 0), torch.cat(pieces_m, 0)
    return render, alpha_img, amb_img, n_isect


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                  near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, backgrounds=None,
                  render_mode="RGB", rasterize_mode="classic", margin=1e-5, row_range=None, absgrad=False):
    """gsplat-shaped entry: colors [N,3] or [C,N,3]; returns renders [C,H,W,D], alphas [C,H,W,1], info.
    ``absgrad=True``: ``info["absgrad"]`` [C,N,2] is filled during backward (see ``composite``)."""
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    C = viewmats.shape[0]
    renders, alphas, ambs = [], [], []
    info = dict(radii=[], means2d=[], depths=[], conics=[], n_isect=[], ambiguous_gauss=[])
    abs_all = torch.zeros(C, means.shape[0], 2, dtype=means.dtype) if absgrad else None
    prs = [project(means, quats, scales, viewmats[c], Ks[c], width, height, eps2d, near_plane, far_plane, radius_clip,
                   margin) for c in range(C)]
    # info["means2d"] is a graph tensor UPSTREAM of the images, as in gsplat (base.py:430 calls retain_grad() on it)
    means2d_all = torch.stack([pr["means2d"] for pr in prs])
    for c in range(C):
        pr = dict(prs[c], means2d=means2d_all[c])
        col = colors[c] if colors.dim() == 3 else colors
        op = opacities
        if rasterize_mode == "antialiased":
            op = opacities * pr["compensations"]
        if render_mode in ("D", "ED"):
            feat = pr["depths"][:, None]
        elif render_mode in ("RGB+D", "RGB+ED"):
            feat = torch.cat([col, pr["depths"][:, None]], dim=-1)
        else:
            feat = col
        bg = None
        if backgrounds is not None:
            bg = backgrounds[c]
            if render_mode in ("RGB+D", "RGB+ED"):
                bg = torch.cat([bg, torch.zeros(1, dtype=bg.dtype)])
        r, a, m, ni = composite(pr["means2d"], pr["conics"], op, feat, pr["depths"], pr["radii"],
                                width, height, bg, margin, row_range, abs_acc=None if abs_all is None else abs_all[c])
        # a rounding-ambiguous Gaussian may gain / lose boundary tiles: flag the tiles in
        # rect(radius + 1) \ rect(radius - 1)
        if pr["ambiguous"].any():
            idx = torch.nonzero(pr["ambiguous"])[:, 0]
            mu = pr["means2d"][idx].detach()
            big = tile_rect(mu, pr["radii"][idx] + 1, width, height)
            small = tile_rect(mu, (pr["radii"][idx] - 1).clamp(min=1), width, height)
            m = m.clone()
            for k in range(idx.numel()):
                for tyi in range(int(big[1][k]), int(big[3][k])):
                    for txi in range(int(big[0][k]), int(big[2][k])):
                        if small[0][k] <= txi < small[2][k] and small[1][k] <= tyi < small[3][k]:
                            continue
                        m[tyi * TILE:(tyi + 1) * TILE, txi * TILE:(txi + 1) * TILE] = True
        if render_mode in ("ED", "RGB+ED"):
            r = torch.cat([r[..., :-1], r[..., -1:] / a.clamp(min=1e-10)[..., None]], dim=-1)
        renders.append(r); alphas.append(a[..., None]); ambs.append(m)
        for k in ("radii", "means2d", "depths", "conics"):
            info[k].append(pr[k])
        info["n_isect"].append(ni)
        info["ambiguous_gauss"].append(pr["ambiguous"])
    for k in ("radii", "depths", "conics", "ambiguous_gauss"):
        info[k] = torch.stack(info[k])
    info["means2d"] = means2d_all
    info["ambiguous"] = torch.stack(ambs)
    info["width"], info["height"] = width, height
    if abs_all is not None:
        info["absgrad"] = abs_all
    return torch.stack(renders), torch.stack(alphas), info
