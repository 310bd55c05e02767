"""Differentiable Gaussian-splat render over the sm_100a kernels: one autograd Function that runs
project(+activations+SH) -> bin/sort -> fused composite(+glue+bilateral) and its backward.

Public entry points

* ``rasterization(...)``         gsplat v1.3.0-shaped call the reference makes at
                                 ``models/trainers/base.py:393-408`` (import seam
                                 ``models/gaussians/basics.py:12``).
* ``spherical_harmonics(...)``   ``gsplat.cuda._wrapper.spherical_harmonics`` as called at
                                 ``models/gaussians/vanilla.py:383-389``.
* ``render_fused(...)``          the whole hot path of one training step in one op: raw
                                 ``VanillaGaussians`` parameters (vanilla.py:122-146 activations and SH
                                 fused), reference glue (clamp, RGB+ED, sky composite:
                                 base.py:414-417, scene_graph.py:287-294) and the multi-scale bilateral
                                 chain (modules.py:505-584, scene_graph.py:112-117) fused in the
                                 composite kernel.

No torch/CPU fallback exists: every op raises if the tensors are not on a CUDA device.
"""
import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ._lib import (COUNTERS_LEN, TILE, BdsError, BilateralDesc, EpilogueDesc, RenderDesc, check, device_scoped, lib, ptr,
                   ptr_array, require_cuda, stream_ptr)

NULL = C.c_void_p(0)

# bench.py sets this to a dict of lists to collect CUDA-event pairs around the C-ABI calls (recorded on the
# launching stream): keys "project_fwd", "bin_sort", "composite_fwd", "composite_bwd", "project_bwd".
KERNEL_EVENTS: Optional[dict] = None


class _timed:
    """Records a CUDA-event pair around one C-ABI call when KERNEL_EVENTS is armed."""

    def __init__(self, key):
        self.key = key

    def __enter__(self):
        if KERNEL_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if KERNEL_EVENTS is not None:
            self.e1.record()
            KERNEL_EVENTS.setdefault(self.key, []).append((self.e0, self.e1))
        return False


@dataclass
class RenderCfg:
    width: int
    height: int
    near_plane: float = 0.01
    far_plane: float = 1e10
    radius_clip: float = 0.0
    eps2d: float = 0.3
    antialiased: bool = False
    row_begin: int = 0
    row_end: int = -1          # -1 = all tile rows of all cameras
    raw_params: bool = False   # log-scales / logit opacities (VanillaGaussians fast path)
    sh_degree: int = -1        # >= 0: colours from features_dc / features_rest
    mode: int = 0              # 0 gsplat outputs, 1 + reference glue, 2 + bilateral chain (full-res guidance)
    channels: int = 3
    expected_depth: bool = False
    bil_sizes: Tuple = ()      # ((grid_X, grid_Y, grid_W), ...)
    absgrad: bool = False
    dense_info: bool = True    # write gsplat-shaped means2d / depths / conics
    splat_capacity: Optional[int] = None
    # multi-GPU (dist.py): exchange the per-splat gradient records instead of all-reducing the dense parameter
    # gradients - the backward all-gathers them over this process group (True = the default group) and runs the
    # projection backward over every rank's records, so the returned Gaussian gradients are already the job's
    exchange_group: object = None
    # "splats": all-gather the per-splat gradient records and run the projection backward over all of them;
    # "compact": all-reduce the Gaussian gradients with the SH part in compact form (3 floats per (camera, Gaussian)
    # instead of 3 K per Gaussian: 11 + 3 C floats instead of 59) inside the backward, then expand it
    exchange_mode: str = "splats"

    def tiles(self):
        return (self.width + TILE - 1) // TILE, (self.height + TILE - 1) // TILE


def _desc(cfg: RenderCfg, N: int, Cn: int, sh_K: int) -> RenderDesc:
    tw, th = cfg.tiles()
    d = RenderDesc()
    d.n_gauss, d.n_cams, d.width, d.height = N, Cn, cfg.width, cfg.height
    d.near_plane, d.far_plane, d.radius_clip, d.eps2d = cfg.near_plane, min(cfg.far_plane, 3.0e38), cfg.radius_clip, cfg.eps2d
    d.antialiased = int(cfg.antialiased)
    d.row_begin = cfg.row_begin
    d.row_end = Cn * th if cfg.row_end < 0 else cfg.row_end
    d.raw_params = int(cfg.raw_params)
    d.sh_degree = cfg.sh_degree
    d.sh_K = sh_K
    return d


def _epilogue(cfg: RenderCfg) -> EpilogueDesc:
    e = EpilogueDesc()
    e.mode, e.channels, e.expected_depth = cfg.mode, cfg.channels, int(cfg.expected_depth)
    if cfg.mode == 2:
        e.bil = BilateralDesc.make(cfg.bil_sizes, None)
    return e


def band_pixel_rows(cfg: RenderCfg, Cn: int) -> Tuple[int, int]:
    """Stacked pixel rows (c*H + y) covered by the band [row_begin, row_end) of tile rows."""
    tw, th = cfg.tiles()
    rb = cfg.row_begin
    re = Cn * th if cfg.row_end < 0 else cfg.row_end

    def row_of(gr):
        c, ty = divmod(gr, th)
        return c * cfg.height + min(ty * TILE, cfg.height)

    return row_of(rb), (row_of(re) if re < Cn * th else Cn * cfg.height)


def check_grid_slots(cfg: RenderCfg, Cn: int, grids: Sequence[Optional[torch.Tensor]]) -> None:
    """Mode 2 runs the bilateral chain on every pixel of every camera that owns a tile row of the band, so such a
    camera needs all its grid slots; ``None`` is only meaningful for cameras outside the band (multi-GPU bands).
    ``grids`` = camera-major flat list ``[c * n_levels + l]``."""
    if cfg.mode != 2:
        return
    n_levels = len(cfg.bil_sizes)
    if len(grids) != Cn * n_levels:
        raise ValueError(f"expected {Cn} x {n_levels} grid slots, got {len(grids)}")
    _, th = cfg.tiles()
    rb = cfg.row_begin
    re = Cn * th if cfg.row_end < 0 else cfg.row_end
    for c in range(Cn):
        if rb < (c + 1) * th and re > c * th:
            missing = [l for l in range(n_levels) if grids[c * n_levels + l] is None]
            if missing:
                raise ValueError(f"camera {c} lies inside the band [{rb}, {re}) of tile rows but has no grid slot for "
                                 f"level(s) {missing}; use mode 1 (grid_slots=None) for a glue-only render")


_PINNED = {}


def _pinned_sync_buffers(dev):
    """Per-device pinned landing buffers of the step's one device -> host read."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _PINNED:
        _PINNED[key] = (torch.zeros(1, dtype=torch.int64).pin_memory(), torch.zeros(2, dtype=torch.int32).pin_memory())
    return _PINNED[key]


class _RenderFn(torch.autograd.Function):
    """Inputs (tensors, may be None): means, quats, scales, opacities, colors, features_dc,
    features_rest, viewmats, Ks, backgrounds, sky, *grids (C * n_levels slots [12,L,GY,GX]).
    Outputs: out_rgb, out_rgb_gauss, out_depth, out_alpha, means2d (dense or empty), radii."""

    @staticmethod
    @device_scoped
    def forward(ctx, cfg: RenderCfg, holder: dict, means, quats, scales, opacities, colors, fdc, frest, viewmats,
                Ks, backgrounds, sky, *grids):
        require_cuda(means, quats, scales, opacities, colors, fdc, frest, viewmats, Ks, backgrounds, sky, *grids)
        dev = means.device
        # unused outputs hand None (not a zero tensor) to backward: v_rgb_gauss / v_depth / v_alpha / the [C,N,2]
        # means2d cotangent are optional in the C ABI
        ctx.set_materialize_grads(False)
        f32 = dict(device=dev, dtype=torch.float32)
        i32 = dict(device=dev, dtype=torch.int32)
        cont = lambda t: None if t is None else t.contiguous().float()  # noqa: E731
        means, quats, scales, opacities, colors, fdc, frest, viewmats, Ks, backgrounds, sky = map(
            cont, (means, quats, scales, opacities, colors, fdc, frest, viewmats, Ks, backgrounds, sky))
        grids = [None if g is None else g.contiguous().float() for g in grids]
        N, Cn = means.shape[0], viewmats.shape[0]
        check_grid_slots(cfg, Cn, grids)
        sh_K = 0
        if cfg.sh_degree >= 0:
            sh_K = 1 + (0 if frest is None else frest.shape[1])
        d = _desc(cfg, N, Cn, sh_K)
        e = _epilogue(cfg)
        tw, th = cfg.tiles()
        n_band_tiles = (d.row_end - d.row_begin) * tw
        r0, r1 = band_pixel_rows(cfg, Cn)
        P = (r1 - r0) * cfg.width
        colors_per_cam = int(colors is not None and colors.dim() == 3)

        radii = torch.zeros(Cn, N, **i32)
        if cfg.dense_info:
            means2d = torch.zeros(Cn, N, 2, **f32)
            depths = torch.zeros(Cn, N, **f32)
            conics = torch.zeros(Cn, N, 3, **f32)
        else:
            means2d = depths = conics = None
        comps = torch.zeros(Cn, N, **f32) if (cfg.antialiased and cfg.dense_info) else None
        tiles_touched = torch.empty(Cn, N, **i32)
        # record index of every (camera, Gaussian), -1 = none: only kept when the gsplat info tensors are exposed (a
        # user cotangent on them must also reach visible Gaussians that own no record: bds_project_bwd_extras)
        slot_of = torch.empty(Cn, N, **i32) if cfg.dense_info else None
        tile_counts = torch.empty(n_band_tiles + 1, **i32)
        counters = torch.zeros(COUNTERS_LEN, **i32)   # [0] records, [1] overflow flag, [2..] queue of very large splats
        cap = cfg.splat_capacity or max(Cn * N, 1)
        splats = torch.empty(cap, 12, **f32)
        st = stream_ptr()
        with _timed("project_fwd"):
            check(lib.bds_project_fwd(C.byref(d), ptr(means), ptr(quats), ptr(scales), ptr(opacities), ptr(colors),
                                      colors_per_cam, ptr(fdc), ptr(frest), ptr(viewmats), ptr(Ks), ptr(radii),
                                      ptr(means2d), ptr(depths), ptr(conics), ptr(comps), ptr(tiles_touched),
                                      ptr(tile_counts), ptr(splats), C.c_int32(cap), ptr(slot_of), ptr(counters), st),
                  "bds_project_fwd")
        stats = torch.zeros(1, device=dev, dtype=torch.int64)
        tile_offsets = torch.empty(n_band_tiles + 1, **i32)
        ws0 = torch.empty(int(lib.bds_bin_count_workspace_bytes(C.byref(d))), device=dev, dtype=torch.uint8)
        check(lib.bds_bin_count(C.byref(d), ptr(tile_counts), ptr(tile_offsets), ptr(stats), ptr(ws0), st), "bds_bin_count")
        # the one host sync of the step: intersection count (sizes the sort) + slot count / overflow flag, copied
        # straight into pinned host memory (no staging kernels, no pageable bounce) and waited for on the stream
        h_stats, h_counters = _pinned_sync_buffers(dev)
        h_stats.copy_(stats, non_blocking=True)
        h_counters.copy_(counters[:2], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        n_isect, n_slots, overflow = int(h_stats[0]), int(h_counters[0]), int(h_counters[1])
        if overflow:
            raise BdsError(f"splat capacity {cap} exceeded ({n_slots} visible splats); raise RenderCfg.splat_capacity")
        ctx.exchange = None
        if cfg.exchange_group is not None:
            from . import dist as D
            if cfg.dense_info:
                raise BdsError("exchange_group needs dense_info=False (the dense densification taps are per rank)")
            grp = None if cfg.exchange_group is True else cfg.exchange_group
            if cfg.exchange_mode == "compact":
                if fdc is None or colors is not None:
                    raise BdsError("exchange_mode='compact' needs the SH fast path (features_dc / features_rest)")
                ctx.exchange = (grp, None, None)
            else:
                counts, total_dev = D.gather_splat_counts(n_slots, dev, grp)
                ctx.exchange = (grp, counts, total_dev)
        sorted_splats = torch.empty(max(n_isect, 1), 12, **f32)
        ws1 = torch.empty(int(lib.bds_bin_sort_workspace_bytes(C.byref(d), C.c_int64(n_isect))), device=dev,
                          dtype=torch.uint8)
        with _timed("bin_sort"):
            check(lib.bds_bin_sort(C.byref(d), C.c_int64(n_isect), C.c_int32(n_slots), ptr(radii), ptr(splats),
                                   ptr(tile_offsets), ptr(sorted_splats), NULL, ptr(ws1), st), "bds_bin_sort")
        del ws1
        ch = cfg.channels if cfg.mode == 0 else 3
        out_rgb = torch.empty(P, ch, **f32)
        out_alpha = torch.empty(P, **f32)
        last_ids = torch.empty(P, **i32)
        out_rgbg = torch.empty(P, 3, **f32) if cfg.mode != 0 else None
        out_depth = torch.empty(P, **f32) if cfg.mode != 0 else None
        ws2 = torch.empty(int(lib.bds_composite_workspace_bytes(C.byref(d), C.byref(e))), device=dev, dtype=torch.uint8)
        with _timed("composite_fwd"):
            check(lib.bds_composite_fwd(C.byref(d), C.byref(e), ptr(sorted_splats), ptr(tile_offsets), ptr(backgrounds),
                                        ptr(sky), ptr_array(grids) if grids else NULL, ptr(out_rgb), ptr(out_rgbg),
                                        ptr(out_depth), ptr(out_alpha), ptr(last_ids), ptr(ws2), st),
                  "bds_composite_fwd")
        ctx.cfg, ctx.d, ctx.e, ctx.holder = cfg, d, e, holder
        ctx.n_slots, ctx.n_isect, ctx.colors_per_cam, ctx.n_grids = n_slots, n_isect, colors_per_cam, len(grids)
        ctx.has = dict(colors=colors is not None, sky=sky is not None, bg=backgrounds is not None)
        ctx.save_for_backward(means, quats, scales, opacities, colors, fdc, frest, viewmats, Ks, backgrounds, sky,
                              splats, counters, sorted_splats, tile_offsets, out_rgbg, out_depth, out_alpha, last_ids,
                              out_rgb if (cfg.mode == 0 and cfg.channels == 4 and cfg.expected_depth) else None,
                              radii, slot_of, *grids)
        holder.update(n_isect=n_isect, n_visible=n_slots, depths=depths, conics=conics, tiles_touched=tiles_touched,
                      tile_offsets=tile_offsets, compensations=comps, sorted_splats=sorted_splats, last_ids=last_ids,
                      # what rasterize_masked() needs to composite again over the same sorted lists
                      cache=dict(cfg=cfg, d=d, e=e, splats=splats, counters=counters, n_slots=n_slots, P=P, N=N, Cn=Cn,
                                 sorted_splats=sorted_splats, tile_offsets=tile_offsets, backgrounds=backgrounds))
        m2d = means2d if means2d is not None else torch.empty(0, **f32)
        ctx.mark_non_differentiable(radii)
        return (out_rgb, out_rgbg if out_rgbg is not None else torch.empty(0, **f32),
                out_depth if out_depth is not None else torch.empty(0, **f32), out_alpha, m2d, radii)

    @staticmethod
    @device_scoped
    def backward(ctx, v_rgb, v_rgbg, v_depth, v_alpha, v_means2d_extra, _v_radii):
        cfg, d, e = ctx.cfg, ctx.d, ctx.e
        (means, quats, scales, opacities, colors, fdc, frest, viewmats, Ks, backgrounds, sky, splats, counters,
         sorted_splats, tile_offsets, out_rgbg, out_depth, out_alpha, last_ids, out_rgb_ed, radii, slot_of,
         *grids) = ctx.saved_tensors
        dev = means.device
        f32 = dict(device=dev, dtype=torch.float32)
        N, Cn = means.shape[0], viewmats.shape[0]
        st = stream_ptr()
        cont = lambda t: None if t is None else t.contiguous().float()  # noqa: E731
        v_rgb, v_rgbg, v_depth, v_alpha = map(cont, (v_rgb, v_rgbg, v_depth, v_alpha))
        if v_rgb is None:
            v_rgb = torch.zeros(out_alpha.shape[0], cfg.channels if cfg.mode == 0 else 3, **f32)
        if cfg.mode == 0:
            v_rgbg = None
            v_depth_in = None
            depth_for_ed = out_rgb_ed[:, 3].contiguous() if out_rgb_ed is not None else None
        else:
            depth_for_ed = out_depth
            v_depth_in = v_depth
            if v_rgbg is not None and v_rgbg.numel() == 0:
                v_rgbg = None
        v_splats = torch.zeros(max(ctx.n_slots, 1), 12, **f32)
        need = ctx.needs_input_grad  # (cfg, holder, means, quats, scales, opac, colors, fdc, frest, viewmats, Ks, bg, sky, *grids)
        v_sky = torch.empty_like(sky) if (sky is not None and need[12]) else None
        v_bg = torch.zeros_like(backgrounds) if (backgrounds is not None and need[11]) else None
        # all grid-slot gradients are carved out of ONE zero-filled buffer (one fill instead of C x levels)
        n_grid = sum(g.numel() for g in grids if g is not None)
        compact = ctx.exchange is not None and cfg.exchange_mode == "compact"
        if compact:
            # ONE buffer for everything the ranks exchange: [Gaussian part (filled by _backward_compact) | grid slots]
            n_gauss_part = N * 10 + opacities.numel() + Cn * N * 3
            big = torch.zeros(n_gauss_part + n_grid, **f32)
            g_flat = big[n_gauss_part:]
        else:
            big = None
            g_flat = torch.zeros(n_grid, **f32)
        v_grids, g_off = [], 0
        for g in grids:
            v_grids.append(None if g is None else g_flat[g_off:g_off + g.numel()].view(g.shape))
            g_off += 0 if g is None else g.numel()
        ws2 = torch.empty(int(lib.bds_composite_workspace_bytes(C.byref(d), C.byref(e))), device=dev, dtype=torch.uint8)
        with _timed("composite_bwd"):
            check(lib.bds_composite_bwd(C.byref(d), C.byref(e), ptr(sorted_splats), NULL, ptr(tile_offsets),
                                        ptr(backgrounds), ptr(sky), ptr_array(grids) if grids else NULL, ptr(out_rgbg),
                                        ptr(depth_for_ed), ptr(out_alpha), ptr(last_ids), ptr(v_rgb), ptr(v_rgbg),
                                        ptr(v_depth_in), ptr(v_alpha), ptr(v_splats), ptr(v_sky),
                                        ptr_array(v_grids) if grids else NULL, ptr(v_bg), ptr(ws2), st),
                  "bds_composite_bwd")
        # every Gaussian-parameter gradient lives in ONE flat zero-filled buffer (one fill, and the
        # multi-GPU step all-reduces it in place: dist.allreduce_grads)
        if ctx.exchange is not None and cfg.exchange_mode == "compact":
            return _RenderFn._backward_compact(ctx, means, quats, scales, opacities, fdc, frest, viewmats, Ks, splats,
                                               counters, v_splats, need, v_bg, v_sky, v_grids, st, big)
        parts = [("means", means), ("quats", quats), ("scales", scales), ("opac", opacities)]
        if colors is not None:
            parts.append(("colors", colors))
        if fdc is not None:
            parts.append(("fdc", fdc))
        if frest is not None:
            parts.append(("frest", frest))
        flat = torch.zeros(sum(t.numel() for _, t in parts), **f32)
        views, off = {}, 0
        for name, t in parts:
            views[name] = flat[off:off + t.numel()].view(t.shape)
            off += t.numel()
        v_means, v_quats, v_scales, v_opac = views["means"], views["quats"], views["scales"], views["opac"]
        v_colors, v_fdc, v_frest = views.get("colors"), views.get("fdc"), views.get("frest")
        ctx.holder["grad_flat"] = flat
        rec_fwd, rec_bwd, n_counter = splats, v_splats, counters
        if ctx.exchange is not None:
            # every rank's (forward record, gradient record) pairs; the projection backward below then produces the
            # gradient of the WHOLE job on every rank (no dense all-reduce of the Gaussian gradients afterwards)
            from . import dist as D
            grp, counts, total_dev = ctx.exchange
            with _timed("exchange"):
                rec_fwd = D.allgather_rows(splats, counts, grp)
                rec_bwd = D.allgather_rows(v_splats, counts, grp)
            n_counter = total_dev
            ctx.holder["grads_are_global"] = True
        v_view = torch.zeros_like(viewmats) if need[9] else None
        want_taps = cfg.dense_info
        v_m2d = torch.zeros(Cn, N, 2, **f32) if want_taps else None
        absg = torch.zeros(Cn, N, 2, **f32) if (want_taps and cfg.absgrad) else None
        extra = None
        if v_means2d_extra is not None and v_means2d_extra.numel() > 0:
            extra = v_means2d_extra.contiguous().float()
        with _timed("project_bwd"):
            # bds_project_bwd launches one thread per possible record of ONE rank (C x N); the gathered records of all
            # ranks normally fit (a splat is repeated only where it reaches two bands), else they go in several calls
            pieces = [(rec_fwd, rec_bwd, n_counter)]
            if ctx.exchange is not None and sum(ctx.exchange[1]) > Cn * N:
                tot, step_n = sum(ctx.exchange[1]), Cn * N
                pieces = [(rec_fwd[o:o + step_n], rec_bwd[o:o + step_n],
                           torch.tensor([min(step_n, tot - o)], device=dev, dtype=torch.int32))
                          for o in range(0, tot, step_n)]
            for pf, pb, pc in pieces:
                check(lib.bds_project_bwd(C.byref(d), ptr(means), ptr(quats), ptr(scales), ptr(opacities), ptr(colors),
                                          ctx.colors_per_cam, ptr(fdc), ptr(frest), ptr(viewmats), ptr(Ks), ptr(pf),
                                          ptr(pc), ptr(pb), ptr(extra), NULL, NULL, ptr(v_means), ptr(v_quats),
                                          ptr(v_scales), ptr(v_opac), ptr(v_colors), ptr(v_fdc), ptr(v_frest), ptr(v_view),
                                          ptr(v_m2d), ptr(absg), st), "bds_project_bwd")
            if extra is not None and slot_of is not None:
                check(lib.bds_project_bwd_extras(C.byref(d), ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks),
                                                 ptr(radii), ptr(slot_of), ptr(extra), NULL, NULL, ptr(v_means),
                                                 ptr(v_quats), ptr(v_scales), ptr(v_view), st), "bds_project_bwd_extras")
        # densification taps (base.py:279-297 reads info["means2d"].grad / .absgrad)
        ref = ctx.holder.get("means2d_ref")
        m2d = ref() if ref is not None else None
        if m2d is not None and want_taps:
            total = v_m2d if extra is None else v_m2d + extra
            if m2d.retains_grad or m2d.is_leaf:
                m2d.grad = total
            if absg is not None:
                m2d.absgrad = absg
        ctx.holder["v_splats"] = v_splats
        return (None, None, v_means, v_quats, v_scales, v_opac, v_colors, v_fdc, v_frest, v_view, None, v_bg, v_sky,
                *v_grids)


def _backward_compact(ctx, means, quats, scales, opacities, fdc, frest, viewmats, Ks, splats, counters, v_splats, need,
                      v_bg, v_sky, v_grids, st, flat):
    """Tail of ``_RenderFn.backward`` for ``exchange_mode="compact"``: projection backward with the SH gradient as one
    colour cotangent per (camera, Gaussian), ONE all-reduce of [means | quats | scales | opacities | that] over the
    exchange group, then the expansion to ``_features_dc`` / ``_features_rest``.  The gradients returned are the job's."""
    import torch.distributed as dist

    cfg, d = ctx.cfg, ctx.d
    dev = means.device
    N, Cn = means.shape[0], viewmats.shape[0]
    f32 = dict(device=dev, dtype=torch.float32)
    sizes = [("means", (N, 3)), ("quats", (N, 4)), ("scales", (N, 3)), ("opac", tuple(opacities.shape)), ("shc", (Cn, N, 3))]
    numel = [int(torch.Size(sh).numel()) for _, sh in sizes]
    views, off = {}, 0      # `flat` = [this Gaussian part | the grid-slot gradients composite_bwd has already written]
    for (name, sh), n_el in zip(sizes, numel):
        views[name] = flat[off:off + n_el].view(sh)
        off += n_el
    v_view = torch.zeros_like(viewmats) if need[9] else None
    with _timed("project_bwd"):
        check(lib.bds_project_bwd_compact_sh(C.byref(d), ptr(means), ptr(quats), ptr(scales), ptr(opacities), ptr(viewmats),
                                             ptr(Ks), ptr(splats), ptr(counters), ptr(v_splats), ptr(views["means"]),
                                             ptr(views["quats"]), ptr(views["scales"]), ptr(views["opac"]),
                                             ptr(views["shc"]), ptr(v_view), st), "bds_project_bwd_compact_sh")
    grp = ctx.exchange[0]
    with _timed("exchange"):
        # pieces of at most 256 MB: NCCL's in-switch reduction (NVLS) was measured slower per byte on a single 928 MB
        # message (8 M Gaussians) than on 232 MB ones
        piece = 64 * 1024 * 1024
        for o in range(0, flat.numel(), piece):
            dist.all_reduce(flat[o:o + piece], op=dist.ReduceOp.SUM, group=grp)
        if v_view is not None:
            dist.all_reduce(v_view, op=dist.ReduceOp.SUM, group=grp)
    v_fdc, v_frest = torch.empty_like(fdc), torch.empty_like(frest)
    with _timed("sh_expand"):
        check(lib.bds_sh_expand_bwd(C.byref(d), ptr(means), ptr(viewmats), ptr(views["shc"]), ptr(v_fdc), ptr(v_frest), st),
              "bds_sh_expand_bwd")
    ctx.holder["grads_are_global"] = True
    ctx.holder["grids_are_global"] = True     # the grid-slot gradients rode in the same all-reduce
    ctx.holder["v_splats"] = v_splats
    return (None, None, views["means"], views["quats"], views["scales"], views["opac"], None, v_fdc, v_frest, v_view, None,
            v_bg, v_sky, *v_grids)


_RenderFn._backward_compact = staticmethod(_backward_compact)


def _run(cfg: RenderCfg, means, quats, scales, opacities, colors, fdc, frest, viewmats, Ks, backgrounds, sky, grids):
    holder: dict = {}
    outs = _RenderFn.apply(cfg, holder, means, quats, scales, opacities, colors, fdc, frest, viewmats, Ks, backgrounds,
                           sky, *grids)
    out_rgb, out_rgbg, out_depth, out_alpha, means2d, radii = outs
    if means2d.numel() > 0:
        holder["means2d_ref"] = weakref.ref(means2d)
    return out_rgb, out_rgbg, out_depth, out_alpha, means2d, radii, holder


def _as_int(v) -> int:
    """width / height arrive as python ints or 0-dim (CUDA) int64 tensors (pixel_source.py:653-654)."""
    return int(v.item()) if torch.is_tensor(v) else int(v)


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane=0.01,
                  far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None, packed=True, tile_size=16,
                  backgrounds=None, render_mode="RGB", sparse_grad=False, absgrad=False, rasterize_mode="classic",
                  channel_chunk=32, distributed=False, covars=None, **unsupported):
    """gsplat v1.3.0-shaped ``rasterization``.  Returns
    ``(render_colors [C,H,W,D], render_alphas [C,H,W,1], info)``.

    ``packed`` only changes the layout of gsplat's intermediate tensors, not the images: it is
    accepted and the (unpacked) info layout the reference trainer reads is always returned.
    """
    if unsupported:
        raise NotImplementedError(f"unsupported rasterization arguments: {sorted(unsupported)}")
    if sparse_grad:
        raise NotImplementedError("sparse_grad=True is not built (reference config: sparse_grad: false)")
    if distributed:
        raise NotImplementedError("gsplat's distributed=True path is not built; see bilateral_driving_b200.dist")
    if covars is not None:
        raise NotImplementedError("covars input is not built (the reference passes quats/scales)")
    if tile_size != 16:
        raise NotImplementedError("tile_size must be 16")
    if render_mode not in ("RGB", "D", "ED", "RGB+D", "RGB+ED"):
        raise ValueError(f"unknown render_mode {render_mode}")
    if rasterize_mode not in ("classic", "antialiased"):
        raise ValueError(f"unknown rasterize_mode {rasterize_mode}")
    if render_mode in ("D", "ED"):
        raise NotImplementedError("depth-only render modes are not built (reference uses RGB and RGB+ED)")
    W, H = _as_int(width), _as_int(height)
    N, Cn = means.shape[0], viewmats.shape[0]
    assert quats.shape == (N, 4) and scales.shape == (N, 3) and opacities.shape == (N,), "bad Gaussian shapes"
    assert viewmats.shape == (Cn, 4, 4) and Ks.shape == (Cn, 3, 3), "bad camera shapes"
    if sh_degree is not None:
        # gsplat would evaluate SH itself and apply clamp(sh + 0.5, min=0); the reference never passes
        # sh_degree (it evaluates colours first, vanilla.py:383-389). The fused SH fast path with the
        # reference's own clamp(+0.5, 0, 1) is render_fused().
        raise NotImplementedError("pass evaluated colours (as the reference does); use render_fused for the SH fast path")
    if colors.shape[-1] != 3:
        raise NotImplementedError("only 3-channel colours are built (reference: rgbs [N,3])")
    fdc = frest = None
    colors_in, deg = colors, -1
    D = 4 if render_mode in ("RGB+D", "RGB+ED") else 3
    bg = backgrounds
    if bg is not None and D == 4:
        bg = torch.cat([bg, torch.zeros(bg.shape[0], 1, device=bg.device, dtype=bg.dtype)], dim=-1)
    cfg = RenderCfg(width=W, height=H, near_plane=float(near_plane), far_plane=float(far_plane),
                    radius_clip=float(radius_clip), eps2d=float(eps2d), antialiased=rasterize_mode == "antialiased",
                    sh_degree=deg, mode=0, channels=D, expected_depth=render_mode == "RGB+ED", absgrad=bool(absgrad))
    out_rgb, _, _, out_alpha, means2d, radii, holder = _run(cfg, means, quats, scales, opacities, colors_in, fdc, frest,
                                                            viewmats, Ks, bg, None, [])
    renders = out_rgb.view(Cn, H, W, D)
    alphas = out_alpha.view(Cn, H, W, 1)
    tw, th = cfg.tiles()
    info = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d, "depths": holder["depths"],
        "conics": holder["conics"], "opacities": opacities.expand(Cn, N) if opacities.dim() == 1 else opacities,
        "tile_width": tw, "tile_height": th, "tiles_per_gauss": holder["tiles_touched"],
        "isect_offsets": holder["tile_offsets"][:-1].view(Cn, th, tw), "width": W, "height": H, "tile_size": 16,
        "n_cameras": Cn, "n_isect": holder["n_isect"], "n_visible": holder["n_visible"],
        "_bds_cache": holder["cache"],   # private: sorted lists for rasterize_masked()
    }
    return renders, alphas, info


@torch.no_grad()
def rasterize_masked(info, gaussian_mask):
    """Re-render the scene of a previous ``rasterization`` call with a per-Gaussian keep mask, reusing its
    projection, binning and depth-sorted tile lists (one composite launch instead of the whole pipeline).

    Replaces the reference's ``render_fn(opacity_mask)`` (``models/trainers/base.py:392-419``), which calls
    ``rasterization`` again with ``opacities * mask`` for every class at eval time
    (``models/trainers/scene_graph.py:296-313``).  ``gaussian_mask`` is a bool/0-1 tensor ``[N]``; a masked-out
    Gaussian is exactly a Gaussian of opacity 0, so the images equal the re-rasterization bit for bit.
    No autograd (the reference calls it under ``torch.no_grad()``).  Returns ``(renders [C,H,W,D], alphas
    [C,H,W,1])`` like ``rasterization``.
    """
    c = info["_bds_cache"] if isinstance(info, dict) and "_bds_cache" in info else info
    cfg, d = c["cfg"], c["d"]
    if not gaussian_mask.is_cuda:
        raise BdsError("bds operators run on CUDA tensors only (no CPU fallback exists)")
    N, Cn, P = c["N"], c["Cn"], c["P"]
    if gaussian_mask.numel() != N:
        raise ValueError(f"gaussian_mask has {gaussian_mask.numel()} entries for {N} Gaussians")
    # the sorted lists do not depend on the epilogue of the render that produced them: a cache of a fused
    # (mode 1 / 2) render is re-composited with the plain gsplat epilogue the reference's render_fn returns
    if cfg.mode == 0:
        e, channels, bg = c["e"], cfg.channels, c["backgrounds"]
    else:
        e, channels, bg = EpilogueDesc(), 4, None
        e.mode, e.channels, e.expected_depth = 0, 4, 1
    dev = gaussian_mask.device
    with torch.cuda.device(dev):
        keep = (gaussian_mask.reshape(-1) != 0).to(torch.uint8).contiguous()
        slot_keep = torch.empty(max(c["n_slots"], 1), device=dev, dtype=torch.uint8)
        st = stream_ptr()
        check(lib.bds_slot_keep(C.byref(d), ptr(c["splats"]), ptr(c["counters"]), C.c_int32(c["n_slots"]), ptr(keep),
                                ptr(slot_keep), st), "bds_slot_keep")
        f32 = dict(device=dev, dtype=torch.float32)
        out_rgb = torch.empty(P, channels, **f32)
        out_alpha = torch.empty(P, **f32)
        last_ids = torch.empty(P, device=dev, dtype=torch.int32)
        check(lib.bds_composite_fwd_masked(C.byref(d), C.byref(e), ptr(c["sorted_splats"]), ptr(c["tile_offsets"]),
                                           ptr(slot_keep), ptr(bg), NULL, NULL, ptr(out_rgb), NULL, NULL,
                                           ptr(out_alpha), ptr(last_ids), NULL, st), "bds_composite_fwd_masked")
    rows = P // cfg.width
    if rows == Cn * cfg.height:
        return out_rgb.view(Cn, cfg.height, cfg.width, channels), out_alpha.view(Cn, cfg.height, cfg.width, 1)
    return out_rgb.view(rows, cfg.width, channels), out_alpha.view(rows, cfg.width, 1)   # a band of stacked pixel rows


class _SHFn(torch.autograd.Function):
    @staticmethod
    @device_scoped
    def forward(ctx, degree, dirs, coeffs):
        require_cuda(dirs, coeffs)
        d2 = dirs.reshape(-1, 3).contiguous().float()
        K = coeffs.shape[-2]
        c2 = coeffs.reshape(-1, K, 3).contiguous().float()
        out = torch.empty(d2.shape[0], 3, device=d2.device, dtype=torch.float32)
        check(lib.bds_sh_fwd(d2.shape[0], degree, K, ptr(d2), ptr(c2), ptr(out), stream_ptr()), "bds_sh_fwd")
        ctx.save_for_backward(d2, c2)
        ctx.degree, ctx.shape_d, ctx.shape_c = degree, dirs.shape, coeffs.shape
        return out.view(*dirs.shape[:-1], 3)

    @staticmethod
    @device_scoped
    def backward(ctx, v_out):
        d2, c2 = ctx.saved_tensors
        K = c2.shape[1]
        v_out = v_out.reshape(-1, 3).contiguous().float()
        v_c = torch.empty_like(c2)
        v_d = torch.empty_like(d2) if ctx.needs_input_grad[1] else None
        check(lib.bds_sh_bwd(d2.shape[0], ctx.degree, K, ptr(d2), ptr(c2), ptr(v_out), ptr(v_c), ptr(v_d), stream_ptr()),
              "bds_sh_bwd")
        return None, (None if v_d is None else v_d.view(ctx.shape_d)), v_c.view(ctx.shape_c)


def spherical_harmonics(degrees_to_use: int, dirs, coeffs, masks=None):
    """gsplat.cuda._wrapper.spherical_harmonics: dirs [...,3], coeffs [...,K,3] -> [...,3]."""
    if not 0 <= degrees_to_use <= 3:
        raise NotImplementedError(f"SH degree {degrees_to_use}: bands 0..3 are built")
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    assert dirs.shape[:-1] == coeffs.shape[:-2], (dirs.shape, coeffs.shape)
    assert dirs.shape[-1] == 3 and coeffs.shape[-1] == 3
    out = _SHFn.apply(int(degrees_to_use), dirs, coeffs)
    if masks is not None:
        out = out * masks[..., None]
    return out


def num_sh_bases(degree: int) -> int:
    """gsplat.cuda_legacy._wrapper.num_sh_bases.  The kernels evaluate bands 0..3 (the reference's ``sh_degree: 3``,
    configs/omnire_ms_bilateral.yaml:69); degree 4, which gsplat would accept, is refused here as everywhere else."""
    if not 0 <= degree <= 3:
        raise NotImplementedError(f"SH degree {degree}: bands 0..3 are built (reference config sh_degree: 3)")
    return (degree + 1) ** 2


def quat_to_rotmat(quat):
    """gsplat.cuda_legacy._torch_impl.quat_to_rotmat (plain torch glue used by the node models)."""
    assert quat.shape[-1] == 4, quat.shape
    w, x, y, z = torch.unbind(torch.nn.functional.normalize(quat, dim=-1), dim=-1)
    mat = torch.stack(
        [1 - 2 * (y ** 2 + z ** 2), 2 * (x * y - w * z), 2 * (x * z + w * y),
         2 * (x * y + w * z), 1 - 2 * (x ** 2 + z ** 2), 2 * (y * z - w * x),
         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x ** 2 + y ** 2)], dim=-1)
    return mat.reshape(quat.shape[:-1] + (3, 3))


def render_fused(params: Dict[str, torch.Tensor], viewmats, Ks, width: int, height: int, sky=None,
                 grid_slots: Optional[Sequence[Sequence[torch.Tensor]]] = None, bil_sizes=(), sh_degree: int = 3,
                 near_plane: float = 0.1, far_plane: float = 1e10, radius_clip: float = 0.0, absgrad: bool = True,
                 row_begin: int = 0, row_end: int = -1, activated: bool = False, dense_info: bool = False,
                 antialiased: bool = False, guidance_factor=None, exchange_group=None, exchange_mode="splats"):
    """One fused pass of the hot path for C cameras.

    ``params``: ``_means [N,3], _scales (log) [N,3], _quats [N,4], _opacities (logit) [N] or [N,1],
    _features_dc [N,3], _features_rest [N,K-1,3]`` (``activated=True``: scales/opacities already
    activated and ``_rgbs [N,3]`` given instead of SH features).
    ``grid_slots[c][l]`` = camera c's grid slot ``[12,L,GY,GX]`` of level l; ``grid_slots[c]`` may be None only for a
    camera outside the band (``grid_slots=None`` altogether = reference glue only, mode 1).
    ``guidance_factor=None``: full-resolution guidance, the bilateral chain runs INSIDE the composite kernel
    (epilogue mode 2).  ``guidance_factor=[4,4,2]`` (the reference's default, modules.py:505): the composite
    kernel stops after the glue (mode 1) and the low-resolution-guidance kernels of ``bilateral.py`` finish the
    chain per camera (the low-res guidance of a pixel needs neighbours outside its tile).  Requires whole
    cameras in the band.
    ``exchange_group`` (multi-GPU band sharding, dist.py): a process group (True = the default one) over which the
    backward all-gathers the per-splat gradient records, so that the Gaussian gradients it returns are the WHOLE
    job's on every rank (``info["grads_are_global"]``) and only the grid gradients remain to be all-reduced.
    Returns dict(rgb, rgb_gaussians, depth, opacity [band pixels ...], radii, info).
    """
    Cn = viewmats.shape[0]
    lowres = grid_slots is not None and guidance_factor is not None
    mode = 2 if (grid_slots is not None and not lowres) else 1
    cfg = RenderCfg(width=width, height=height, near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip,
                    antialiased=antialiased, row_begin=row_begin, row_end=row_end, raw_params=not activated,
                    sh_degree=-1 if activated else sh_degree, mode=mode, channels=4, expected_depth=True,
                    bil_sizes=tuple(tuple(s) for s in bil_sizes), absgrad=absgrad, dense_info=dense_info,
                    exchange_group=exchange_group, exchange_mode=exchange_mode)
    grids: List[Optional[torch.Tensor]] = []
    if mode == 2:
        assert len(grid_slots) == Cn and all(len(g) == len(bil_sizes) for g in grid_slots if g is not None)
        for c in range(Cn):
            grids += list(grid_slots[c]) if grid_slots[c] is not None else [None] * len(bil_sizes)
    opac = params["_opacities"].reshape(-1)
    if activated:
        out = _run(cfg, params["_means"], params["_quats"], params["_scales"], opac, params["_rgbs"], None, None,
                   viewmats, Ks, None, sky, grids)
    else:
        out = _run(cfg, params["_means"], params["_quats"], params["_scales"], opac, None, params["_features_dc"],
                   params["_features_rest"], viewmats, Ks, None, sky, grids)
    out_rgb, out_rgbg, out_depth, out_alpha, means2d, radii, holder = out
    r0, r1 = band_pixel_rows(cfg, Cn)
    rows = r1 - r0
    if lowres:
        from .bilateral import multiscale_bilateral

        if r0 % height or r1 % height:
            raise NotImplementedError("low-resolution guidance needs whole cameras in the band (no halo exchange built)")
        pre = out_rgb.view(rows // height, height, width, 3)
        c0 = r0 // height
        out_rgb = torch.cat([multiscale_bilateral(pre[i], grid_slots[c0 + i], bil_sizes, guidance_factor)
                             for i in range(pre.shape[0])]).reshape(-1, 3)
    return dict(rgb=out_rgb.view(rows, width, 3), rgb_gaussians=out_rgbg.view(rows, width, 3),
                depth=out_depth.view(rows, width, 1), opacity=out_alpha.view(rows, width, 1), radii=radii,
                means2d=means2d, info=holder, pixel_rows=(r0, r1))


class _PhotoLossFn(torch.autograd.Function):
    """mean((rgb-gt)^2) + lambda_d*mean(depth) + lambda_a*mean(alpha) in one kernel that also
    writes the cotangents (the benchmark step's loss)."""

    @staticmethod
    @device_scoped
    def forward(ctx, rgb, gt, depth, alpha, lambda_d, lambda_a, count, unit_cotangent):
        require_cuda(rgb, gt, depth, alpha)
        rgb_c, gt_c = rgb.contiguous(), gt.contiguous()
        depth_c = None if depth is None else depth.contiguous()
        alpha_c = None if alpha is None else alpha.contiguous()
        n = rgb_c.numel() // 3
        loss = torch.zeros((), device=rgb.device, dtype=torch.float32)
        v_rgb = torch.empty_like(rgb_c)
        v_d = torch.empty_like(depth_c) if depth_c is not None else None
        v_a = torch.empty_like(alpha_c) if alpha_c is not None else None
        check(lib.bds_loss_fwd_bwd(C.c_int64(n), ptr(rgb_c), ptr(gt_c), ptr(depth_c), ptr(alpha_c), C.c_float(lambda_d),
                                   C.c_float(lambda_a), C.c_float(1.0 / count), ptr(loss), ptr(v_rgb), ptr(v_d), ptr(v_a),
                                   stream_ptr()), "bds_loss_fwd_bwd")
        ctx.save_for_backward(v_rgb, v_d, v_a)
        ctx.unit = unit_cotangent
        return loss

    @staticmethod
    def backward(ctx, v_loss):
        v_rgb, v_d, v_a = ctx.saved_tensors
        if ctx.unit:  # the loss enters the objective with weight 1: the stored cotangents are final
            return v_rgb, None, v_d, v_a, None, None, None, None
        s = v_loss
        return v_rgb * s, None, (None if v_d is None else v_d * s), (None if v_a is None else v_a * s), None, None, None, None


def photometric_loss(rgb, gt, depth=None, alpha=None, lambda_d=0.0, lambda_a=0.0, count=None, unit_cotangent=False):
    """``count`` = number of pixels of the WHOLE job (so that band-sharded ranks sum to the global mean).
    ``unit_cotangent=True`` promises that the returned loss is added to the objective with weight 1 (scale it
    through the lambdas instead); the backward then returns the cotangents the kernel already wrote instead of
    re-scaling three full images."""
    n = rgb.numel() // 3
    return _PhotoLossFn.apply(rgb, gt, depth, alpha, float(lambda_d), float(lambda_a), float(count or n),
                              bool(unit_cotangent))
