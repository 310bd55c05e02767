"""Host-side mirror of the reference's bilateral-grid interface over the sm_100a kernels.

Same names, constructor arguments, state-dict keys and error behaviour as the reference
(paths relative to /root/reference/project) so that released checkpoints load with
``strict=True`` and the YAML dotted path ``model.Affine.type`` can point here:

* ``BilateralGrid``, ``slice``, ``total_variation_loss``, ``color_affine_transform``
  - ``bilateral/lib_bilagrid.py:135-230, 256-368``
* ``BilateralAffineTransform`` - ``models/modules.py:275-351``
* ``MultiScaleBilateralAffineTransform`` - ``models/modules.py:422-593``

plus ``multiscale_bilateral(rgb, slots, sizes, factors)``: the fused slice + sequential apply
(``models/trainers/scene_graph.py:112-117``) that never materialises the 12-channel full-resolution
affine fields.  All arithmetic runs in ``libbds_b200.so``; there is no torch fallback.
"""
import ctypes as C
from typing import Sequence

import torch
from torch import nn

from ._lib import BilateralDesc, check, device_scoped, lib, ptr, ptr_array, require_cuda, stream_ptr


def _workspace(desc, H, W, device):
    nbytes = lib.bds_bilateral_workspace_bytes(C.byref(desc), H, W)
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class _MSBilateralFn(torch.autograd.Function):
    """rgb [H,W,3] + per-level grid slots [12,L,GY,GX] -> rgb_out [H,W,3] (+ affine fields)."""

    @staticmethod
    @device_scoped
    def forward(ctx, rgb, sizes, factors, want_affine, *slots):
        require_cuda(rgb, *slots)
        ctx.set_materialize_grads(False)   # unused affine-field outputs hand None to backward (100 MB per level at 1080p)
        if rgb.dim() != 3 or rgb.shape[-1] != 3:
            raise ValueError(f"rgb must be [H,W,3], got {tuple(rgb.shape)}")
        H, W, _ = rgb.shape
        desc = BilateralDesc.make(sizes, factors)
        rgb_c = rgb.contiguous().float()
        slots_c = [s.contiguous().float() for s in slots]
        for s, (gx, gy, gl) in zip(slots_c, sizes):
            if tuple(s.shape) != (12, gl, gy, gx):
                raise ValueError(f"grid slot shape {tuple(s.shape)} != (12,{gl},{gy},{gx})")
        out = torch.empty_like(rgb_c)
        affines = [torch.empty(H, W, 12, device=rgb.device, dtype=torch.float32) for _ in slots] if want_affine else []
        ws = _workspace(desc, H, W, rgb.device)
        check(lib.bds_bilateral_fwd(C.byref(desc), H, W, ptr(rgb_c), ptr_array(slots_c), ptr(out),
                                    ptr_array(affines) if want_affine else C.c_void_p(0), ptr(ws), stream_ptr()),
              "bds_bilateral_fwd")
        ctx.desc = desc
        ctx.n_levels = len(slots)
        ctx.want_affine = want_affine
        ctx.save_for_backward(rgb_c, *slots_c)
        if want_affine:
            return (out, *[a.view(1, H, W, 3, 4) for a in affines])
        return out

    @staticmethod
    @device_scoped
    def backward(ctx, v_out, *v_affines):
        rgb_c, *slots_c = ctx.saved_tensors
        H, W, _ = rgb_c.shape
        v_out = v_out.contiguous().float() if v_out is not None else torch.zeros_like(rgb_c)
        v_aff = None
        if ctx.want_affine and any(v is not None for v in v_affines):
            v_aff = [None if v is None else v.contiguous().float() for v in v_affines]
        v_rgb = torch.empty_like(rgb_c)
        v_slots = [torch.zeros_like(s) for s in slots_c]
        ws = _workspace(ctx.desc, H, W, rgb_c.device)
        check(lib.bds_bilateral_bwd(C.byref(ctx.desc), H, W, ptr(rgb_c), ptr_array(slots_c), ptr(v_out),
                                    ptr_array(v_aff) if v_aff is not None else C.c_void_p(0), ptr(v_rgb),
                                    ptr_array(v_slots), ptr(ws), stream_ptr()),
              "bds_bilateral_bwd")
        return (v_rgb, None, None, None, *v_slots)


def multiscale_bilateral(rgb, slots: Sequence[torch.Tensor], sizes, factors=(4, 4, 2), return_affine=False):
    """Fused multi-scale slice + sequential apply.  ``slots[l]`` is the image's grid ``[12,L,GY,GX]``
    (``grids[idx]``, or the neighbour average at test time); ``sizes[l] = (grid_X, grid_Y, grid_W)``;
    ``factors`` = the reference's ``guidance_factor`` (None = full-resolution guidance)."""
    sizes = tuple(tuple(int(v) for v in s) for s in sizes)
    factors = None if factors is None else tuple(int(f) for f in factors)
    res = _MSBilateralFn.apply(rgb, sizes, factors, bool(return_affine), *slots)
    if return_affine:
        return res[0], list(res[1:])
    return res


class _SliceFn(torch.autograd.Function):
    """Generic per-sample slice: xy [n,2], rgb [n,3], grid [12,L,GY,GX] -> affine [n,12]."""

    @staticmethod
    @device_scoped
    def forward(ctx, grid, xy, rgb):
        require_cuda(grid, xy, rgb)
        grid_c, xy_c, rgb_c = grid.contiguous().float(), xy.contiguous().float(), rgb.contiguous().float()
        _, L, GY, GX = grid_c.shape
        n = xy_c.shape[0]
        out = torch.empty(n, 12, device=grid.device, dtype=torch.float32)
        check(lib.bds_bilagrid_slice_fwd(ptr(grid_c), L, GY, GX, n, ptr(xy_c), ptr(rgb_c), ptr(out), stream_ptr()),
              "bds_bilagrid_slice_fwd")
        ctx.save_for_backward(grid_c, xy_c, rgb_c)
        return out

    @staticmethod
    @device_scoped
    def backward(ctx, v_aff):
        grid_c, xy_c, rgb_c = ctx.saved_tensors
        _, L, GY, GX = grid_c.shape
        n = xy_c.shape[0]
        v_grid = torch.zeros_like(grid_c)
        v_rgb = torch.zeros_like(rgb_c)
        check(lib.bds_bilagrid_slice_bwd(ptr(grid_c), L, GY, GX, n, ptr(xy_c), ptr(rgb_c),
                                         ptr(v_aff.contiguous().float()), ptr(v_grid), ptr(v_rgb), stream_ptr()),
              "bds_bilagrid_slice_bwd")
        return v_grid, None, v_rgb  # xy carries no gradient in the reference's use (lattice constants)


class _TVFn(torch.autograd.Function):
    """sum_l weight_l * tv(grids_l) for any number of levels in ONE launch, which also writes the gradients (unit
    cotangent); the backward only scales them by the upstream cotangent (which stays on the device)."""

    @staticmethod
    @device_scoped
    def forward(ctx, weights, *grids):
        require_cuda(*grids)
        gs = [g.contiguous().float() for g in grids]
        n = len(gs)
        loss = torch.zeros((), device=gs[0].device, dtype=torch.float32)
        need = [g.requires_grad for g in grids]
        v_gs = [torch.zeros_like(g) if nd else None for g, nd in zip(gs, need)]
        arr = lambda vals, ct: (ct * n)(*vals)  # noqa: E731
        check(lib.bds_tv_levels_fwd_bwd(n, ptr_array(gs), arr([g.shape[0] for g in gs], C.c_int),
                                        arr([g.shape[2] for g in gs], C.c_int), arr([g.shape[3] for g in gs], C.c_int),
                                        arr([g.shape[4] for g in gs], C.c_int), arr([float(w) for w in weights], C.c_float),
                                        C.c_float(1.0), ptr(loss), ptr_array(v_gs), stream_ptr()), "bds_tv_levels_fwd_bwd")
        ctx.v_gs = v_gs
        return loss

    @staticmethod
    @device_scoped
    def backward(ctx, v_loss):
        live = [v for v in ctx.v_gs if v is not None]
        if live:
            torch._foreach_mul_(live, v_loss.to(live[0].dtype))
        return (None, *ctx.v_gs)


def total_variation_loss(x, weight: float = 1.0):
    """lib_bilagrid.py:152-168 for x of shape (B, 12, L, GY, GX)."""
    if x.dim() != 5 or x.shape[1] != 12:
        raise ValueError("total_variation_loss expects a (B,12,L,H,W) bilateral grid tensor")
    return _TVFn.apply((float(weight),), x)


def total_variation_loss_levels(grids, weights):
    """sum_l weights[l] * total_variation_loss(grids[l]) in one launch (modules.py:466-472)."""
    grids = list(grids)
    for x in grids:
        if x.dim() != 5 or x.shape[1] != 12:
            raise ValueError("total_variation_loss expects (B,12,L,H,W) bilateral grid tensors")
    return _TVFn.apply(tuple(float(w) for w in weights), *grids)


def color_affine_transform(affine_mats, rgb):
    """lib_bilagrid.py:135-145 (plain torch glue; not on the fused path)."""
    return torch.matmul(affine_mats[..., :3], rgb.unsqueeze(-1)).squeeze(-1) + affine_mats[..., 3]


class BilateralGrid(nn.Module):
    """lib_bilagrid.py:256-368.  Parameter ``grids`` (N,12,L,H,W), buffer ``rgb2gray_weight`` [1,3]."""

    def __init__(self, num, grid_X=16, grid_Y=16, grid_W=8, mode="bilinear"):
        super().__init__()
        if mode != "bilinear":
            raise NotImplementedError("only the reference's default mode='bilinear' is built")
        self.grid_width, self.grid_height, self.grid_guidance, self.mode = grid_X, grid_Y, grid_W, mode
        eye = torch.tensor([1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0])
        grid = eye.view(1, 12, 1, 1, 1).expand(num, 12, grid_W, grid_Y, grid_X).contiguous()
        self.grids = nn.Parameter(grid)
        self.register_buffer("rgb2gray_weight", torch.Tensor([[0.299, 0.587, 0.114]]))

    @property
    def size_xyl(self):
        return (self.grid_width, self.grid_height, self.grid_guidance)

    def tv_loss(self):
        return total_variation_loss(self.grids)

    def forward(self, grid_xy, rgb, idx=None):
        """Slices with arbitrary xy in [0,1]; 2-D..5-D inputs as in the reference."""
        nd = grid_xy.dim()
        if rgb.dim() != nd:
            raise AssertionError("grid_xy and rgb must have the same number of dims")
        if 1 < nd < 5:
            for _ in range(5 - nd):
                grid_xy, rgb = grid_xy.unsqueeze(1), rgb.unsqueeze(1)
            assert idx is not None
        elif nd != 5:
            raise ValueError("Bilateral grid slicing only takes either 2D, 3D, 4D and 5D inputs")
        grids = self.grids if idx is None else self.grids[idx]
        if grids.dim() == 4:
            grids = grids[None]
        assert grids.shape[0] == grid_xy.shape[0]
        outs = []
        for b in range(grids.shape[0]):
            a = _SliceFn.apply(grids[b], grid_xy[b].reshape(-1, 2), rgb[b].reshape(-1, 3))
            outs.append(a.reshape(*grid_xy.shape[1:-1], 3, 4))
        affine = torch.stack(outs, 0)
        for _ in range(5 - nd):
            affine = affine.squeeze(1)
        return affine


def slice(bil_grids, xy, rgb, grid_idx):  # noqa: A001 - name mirrors the reference
    """lib_bilagrid.py:171-230."""
    sh_ = rgb.shape
    grid_idx_unique = torch.unique(grid_idx)
    if len(grid_idx_unique) == 1:
        grid_idx = grid_idx_unique
        xy, rgb = xy.unsqueeze(0), rgb.unsqueeze(0)
    else:
        if grid_idx.dim() == 4:
            grid_idx = grid_idx[:, 0, 0, 0]
        elif grid_idx.dim() == 3:
            grid_idx = grid_idx[:, 0, 0]
        elif grid_idx.dim() == 2:
            grid_idx = grid_idx[:, 0]
        else:
            raise ValueError("The input to bilateral grid slicing is not supported yet.")
    affine_mats = bil_grids(xy, rgb, grid_idx)
    rgb = color_affine_transform(affine_mats, rgb)
    return {
        "rgb": rgb.reshape(*sh_),
        "rgb_affine_mats": affine_mats.reshape(*sh_[:-1], affine_mats.shape[-2], affine_mats.shape[-1]),
    }


def _cam_index(image_infos) -> int:
    """Host-side image index (modules.py:507 does ``int(image_infos["img_idx"][0][0])``: a device->host sync)."""
    assert "img_idx" in image_infos
    if "img_idx_host" in image_infos:  # optional: a caller that knows the index on the host avoids the sync
        return int(image_infos["img_idx_host"])
    return int(image_infos["img_idx"][0][0])


def _train_slot(grids, image_infos):
    """``grids[img_idx]`` of the training branch WITHOUT the device->host sync of modules.py:507: the slot is
    selected on the device from the index tensor itself.  Same values and the same dense, one-hot-row gradient as
    the reference's advanced indexing (lib_bilagrid.py:346-348)."""
    if "img_idx_host" in image_infos:
        return grids[int(image_infos["img_idx_host"])]
    idx = image_infos["img_idx"]
    if not torch.is_tensor(idx):
        return grids[int(idx)]
    return grids.index_select(0, idx.reshape(-1)[:1].to(device=grids.device, dtype=torch.long))[0]


def chain_inverse_apply(affines, x):
    """Closed form of what modules.py:474-492 gets from a batched 4x4 ``torch.inverse``: the per-pixel chain
    ``x -> A_n(...A_1(A_0 x))`` of 3x4 affines is itself a 3x4 affine [R | t] (R = R_n...R_0); its inverse maps
    ``y -> R^-1 (y - t)`` with R^-1 = adj(R) / det(R).  ``affines``: list of [..., 3, 4]; ``x``: [..., 3]."""
    R = affines[0][..., :3, :3]
    t = affines[0][..., :3, 3]
    for A in affines[1:]:
        Rl = A[..., :3, :3]
        t = (Rl @ t[..., None])[..., 0] + A[..., :3, 3]
        R = Rl @ R
    a, b, c = R[..., 0, 0], R[..., 0, 1], R[..., 0, 2]
    d, e, f = R[..., 1, 0], R[..., 1, 1], R[..., 1, 2]
    g, h, i = R[..., 2, 0], R[..., 2, 1], R[..., 2, 2]
    c00, c01, c02 = e * i - f * h, c * h - b * i, b * f - c * e
    c10, c11, c12 = f * g - d * i, a * i - c * g, c * d - a * f
    c20, c21, c22 = d * h - e * g, b * g - a * h, a * e - b * d
    det = a * c00 + b * c10 + c * c20
    y = x - t
    y0, y1, y2 = y[..., 0], y[..., 1], y[..., 2]
    out = torch.stack([c00 * y0 + c01 * y1 + c02 * y2, c10 * y0 + c11 * y1 + c12 * y2,
                       c20 * y0 + c21 * y1 + c22 * y2], dim=-1)
    return out / det[..., None]


class BilateralAffineTransform(nn.Module):
    """models/modules.py:275-351 (single grid, full-resolution guidance)."""

    def __init__(self, class_name, n, grid_X, grid_Y, grid_W, device="cuda"):
        super().__init__()
        self.bil_grids = BilateralGrid(num=n, grid_X=grid_X, grid_Y=grid_Y, grid_W=grid_W)
        self.register_buffer("rgb2gray_weight", torch.Tensor([0.299, 0.587, 0.114]))
        self.class_prefix = class_name + "#"
        self.device = device
        self.in_test_set = False

    def tv_loss(self):
        return total_variation_loss(self.bil_grids.grids)

    def _slots(self, image_infos):
        if not self.in_test_set:
            return [_train_slot(self.bil_grids.grids, image_infos)]
        near = self.training_indices_for_test[_cam_index(image_infos)]
        return [torch.stack([self.bil_grids.grids[i] for i in near]).mean(0)]

    def level_sizes(self):
        return [self.bil_grids.size_xyl]

    def forward(self, rgb, image_infos):
        _, aff = multiscale_bilateral(rgb, self._slots(image_infos), self.level_sizes(), None, True)
        return aff[0]

    def transform(self, rgb, image_infos):
        """Fused forward + apply of scene_graph.py:95-98."""
        return multiscale_bilateral(rgb, self._slots(image_infos), self.level_sizes(), None)

    def get_param_groups(self):
        return {self.class_prefix + "all": self.bil_grids.parameters()}


def affine_to_homogeneous_batch(affine_matrices):
    """models/modules.py:352-358."""
    b, h, w, _, _ = affine_matrices.shape
    hom = torch.zeros((b, h, w, 4, 4), device=affine_matrices.device)
    hom[:, :, :, :3, :] = affine_matrices
    hom[:, :, :, 3, 3] = 1
    return hom


class MultiScaleBilateralAffineTransform(nn.Module):
    """models/modules.py:422-593.  ``forward`` keeps the reference's return value (a list of
    [1,H,W,3,4] affine fields, also stored in ``save_matrix``); ``transform`` is the fused
    slice + apply the drop-in trainer uses instead (no 100 MB-per-level fields)."""

    def __init__(self, class_name, n, grid, device="cuda"):
        super().__init__()
        self.grid_size = grid
        self.tv_weight = []
        for i in range(len(self.grid_size)):
            setattr(self, f"bil_grids{i}", BilateralGrid(num=n, grid_X=grid[i][0], grid_Y=grid[i][1], grid_W=grid[i][2]))
            self.tv_weight.append(0.5 * (grid[i][0] * grid[i][1] * grid[i][2]) ** 0.5)
        self.register_buffer("rgb2gray_weight", torch.Tensor([0.299, 0.587, 0.114]))
        self.class_prefix = class_name + "#"
        self.device = device
        self.in_test_set = False
        self.save_matrix = None
        self._cycle_ctx = None

    def _levels(self):
        return [getattr(self, f"bil_grids{i}") for i in range(len(self.grid_size))]

    def tv_loss(self):
        return total_variation_loss_levels([bg.grids for bg in self._levels()], self.tv_weight)

    def _slots(self, image_infos):
        assert "img_idx" in image_infos
        if not self.in_test_set:
            return [_train_slot(bg.grids, image_infos) for bg in self._levels()]
        near = self.training_indices_for_test[_cam_index(image_infos)]
        # the mean over neighbour slices equals the slice of the mean grid (slicing is linear in the
        # grid values at fixed coordinates): modules.py:523-538
        return [torch.stack([bg.grids[i] for i in near]).mean(0) for bg in self._levels()]

    def level_sizes(self):
        return [bg.size_xyl for bg in self._levels()]

    def forward(self, rgb, image_infos, guidance_factor=[4, 4, 2]):  # noqa: B006 - reference signature
        _, out_list = multiscale_bilateral(rgb, self._slots(image_infos), self.level_sizes(), guidance_factor, True)
        self.save_matrix = out_list
        self._cycle_ctx = None
        return out_list

    def transform(self, rgb, image_infos, guidance_factor=[4, 4, 2]):  # noqa: B006
        """Fused slice + sequential apply (scene_graph.py:112-117): no affine fields are materialised.  What
        ``inverse_loss`` would need of them is remembered as (rgb, slots, guidance_factor) and only evaluated if
        the cycle loss is actually asked for."""
        slots = self._slots(image_infos)
        self.remember_for_inverse_loss(rgb, slots, guidance_factor)
        return multiscale_bilateral(rgb, slots, self.level_sizes(), guidance_factor)

    def remember_for_inverse_loss(self, rgb, slots, guidance_factor):
        self.save_matrix = None
        self._cycle_ctx = (rgb, list(slots), guidance_factor)

    def _affine_fields(self):
        if self.save_matrix is not None:
            return self.save_matrix
        if getattr(self, "_cycle_ctx", None) is None:
            raise RuntimeError("inverse_loss called before forward/transform (modules.py:474 reads save_matrix)")
        rgb, slots, gf = self._cycle_ctx
        _, fields = multiscale_bilateral(rgb, slots, self.level_sizes(), gf, True)
        return fields

    def inverse_loss(self, gt, render):
        """modules.py:474-492: mean |chain^-1(gt) - render| - the per-pixel 4x4 ``torch.inverse`` of the reference
        replaced by the closed-form inverse of the composed 3x4 chain (``chain_inverse_apply``)."""
        fields = [a.reshape(gt.shape[0], gt.shape[1], 3, 4) for a in self._affine_fields()]
        return torch.abs(chain_inverse_apply(fields, gt) - render).mean()

    def get_param_groups(self):
        return {f"{self.class_prefix}grid{i}": bg.parameters() for i, bg in enumerate(self._levels())}
