"""Shared by the CPU and GPU trainer tests: builds the reference's UNMODIFIED ``MultiTrainer``
(``models/trainers/scene_graph.py``) and the drop-in ``FusedMultiTrainer`` on the same small synthetic scene, with the
config shape of ``configs/omnire_ms_bilateral.yaml`` (Background = ``VanillaGaussians``, a sky model, the multi-scale
bilateral ``Affine``, ``CamPose`` = the reference's own ``CameraOptModule``).

The reference tree comes from ``/root/reference/project`` or from the byte-for-byte copies under ``oracle/_ref``
(``oracle/build_ref.py``); missing third-party packages are the stand-ins of ``oracle/ref_stubs.py``; ``gsplat`` is
``oracle/gsplat_seam`` (CPU tensors -> the CPU oracle, CUDA tensors -> the product)."""
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.dirname(os.path.abspath(__file__))):   # import_str("trainer_harness.GradientSky") must resolve
    if _p not in sys.path:
        sys.path.insert(0, _p)
GRID = [[2, 2, 1], [4, 4, 2], [8, 8, 4]]                      # omnire_ms_bilateral.yaml:249
REF_MS = "models.modules.MultiScaleBilateralAffineTransform"
OUR_MS = "bilateral_driving_b200.bilateral.MultiScaleBilateralAffineTransform"
REF_SINGLE = "models.modules.BilateralAffineTransform"
OUR_SINGLE = "bilateral_driving_b200.bilateral.BilateralAffineTransform"


class GradientSky(nn.Module):
    """Stand-in for ``models.modules.EnvLight`` (which needs nvdiffrast): a trainable colour times a vertical ramp.
    Same interface: ctor (class_name, n, device, **params), ``forward(image_infos) -> [H,W,3]``, ``get_param_groups``."""

    def __init__(self, class_name, n, device="cpu", **_params):
        super().__init__()
        self.class_prefix = class_name + "#"
        self.color = nn.Parameter(torch.tensor([0.2, 0.5, 0.9]))

    def forward(self, image_infos):
        H, W = image_infos["pixels"].shape[:2]
        ramp = torch.linspace(0.3, 1.0, H, device=self.color.device)[:, None, None]
        return (torch.sigmoid(self.color)[None, None, :] * ramp).expand(H, W, 3)

    def get_param_groups(self):
        return {self.class_prefix + "all": self.parameters()}


def load_scene_graph():
    from oracle.ref_loader import load_reference_trainers

    return load_reference_trainers(os.path.join(ROOT, "oracle", "gsplat_seam"))


def make_cfg(affine_type, w1=0.0, grid=None, single=False, background_type="models.gaussians.VanillaGaussians"):
    from oracle.ref_stubs import Cfg

    lr = dict(lr=6.0e-4, lr_final=3e-5, warmup_steps=1000, lr_pre_warmup=0)
    if single:
        affine = dict(type=affine_type, params=dict(grid_X=4, grid_Y=4, grid_W=2), optim=dict(all=dict(lr=6e-4)))
    else:
        affine = dict(type=affine_type, params=dict(grid=grid or GRID), optim=dict(grid0=lr, grid1=lr, grid2=lr))
    return Cfg(dict(
        trainer=dict(
            type="models.trainers.MultiTrainer",
            optim=dict(num_iters=30000, use_grad_scaler=False, cache_buffer_freq=-1),
            render=dict(near_plane=0.1, far_plane=1e10, antialiased=False, packed=False, absgrad=True, sparse_grad=False,
                        batch_size=1),
            losses=dict(rgb=dict(w=0.8), ssim=dict(w=0.2), mask=dict(w=0.05, opacity_loss_type="bce"),
                        depth=dict(w=0.01, inverse_depth=False, normalize=False, loss_type="l1"),
                        affine=dict(w=0.01, w1=w1)),
            res_schedule=dict(double_steps=250, downscale_times=2),
            gaussian_optim_general_cfg=dict(
                xyz=dict(lr=1.6e-4, lr_final=1.6e-6, scale_factor="scene_radius"), sh_dc=dict(lr=0.0025),
                sh_rest=dict(lr=0.000125), opacity=dict(lr=0.05), scaling=dict(lr=0.005), rotation=dict(lr=0.001)),
            gaussian_ctrl_general_cfg=dict(
                warmup_steps=500, reset_alpha_interval=3000, refine_interval=100, sh_degree_interval=1000,
                n_split_samples=2, reset_alpha_value=0.01, densify_grad_thresh=0.0005, densify_size_thresh=0.003,
                cull_alpha_thresh=0.005, cull_scale_thresh=0.5, cull_screen_size=0.15, split_screen_size=0.05,
                stop_screen_size_at=4000, stop_split_at=15000, sh_degree=3)),
        model=dict(
            Background=dict(type=background_type, reg=dict(sharp_shape_reg=None)),
            Sky=dict(type="trainer_harness.GradientSky", params=dict(), optim=dict(all=dict(lr=0.01))),
            Affine=affine,
            CamPose=dict(type="models.modules.CameraOptModule", optim=dict(all=dict(lr=1e-5, weight_decay=1e-6))))))


N_IMAGES = 4


def build_trainer(cls, cfg, device):
    kw = dict(cfg.trainer)
    trainer = cls(**kw, num_timesteps=N_IMAGES, model_config=cfg.model, num_train_images=N_IMAGES,
                  num_full_images=N_IMAGES, test_set_indices=[],
                  scene_aabb=torch.tensor([[-12.0, -12.0, -2.0], [12.0, 12.0, 6.0]]), device=device)
    return trainer


def init_scene(trainer, device, n=1500, sh_step=3001):   # 3001: SH degree 3, and not a refinement step (% 100)
    """Same synthetic Gaussians as the rasteriser golden scene; grids perturbed off identity; CamPose off zero."""
    from bilateral_driving_b200 import synthetic as S

    p = S.make_gaussians(n, extent=10.0, scale_mean=0.12)
    p["_means"][:, 2] = p["_means"][:, 2] * 0.5
    bg = trainer.models["Background"]
    for k, v in p.items():
        v = v.reshape(-1, 1) if k == "_opacities" else v
        setattr(bg, k, nn.Parameter(v.clone().to(device)))
    bg.step = sh_step                                   # SH degree = min(step // 1000, 3)  (vanilla.py:387)
    g = torch.Generator().manual_seed(99)
    aff = trainer.models["Affine"]
    for name, prm in aff.named_parameters():
        prm.data += 0.05 * torch.randn(prm.shape, generator=g).to(device)
    cam = trainer.models["CamPose"]
    cam.embeds.weight.data = (1e-3 * torch.randn(cam.embeds.weight.shape, generator=g)).to(device)
    trainer.step = sh_step


def copy_models(src, dst, device):
    """Same parameters in both arms.  The state-dict keys of the two Affine modules are identical (the
    checkpoint contract), so this is a strict load."""
    for name, m in src.models.items():
        sd = {k: v.detach().clone().to(device) for k, v in m.state_dict().items()}
        if name == "Background":
            dst.models[name].load_state_dict(sd)      # VanillaGaussians re-allocates to the checkpoint's N
            dst.models[name].step = src.models[name].step
        else:
            dst.models[name].load_state_dict(sd, strict=True)
    dst.step = src.step


def make_batch(device, H=56, W=88, img_idx=1, seed=5):
    from bilateral_driving_b200 import synthetic as S

    g = torch.Generator().manual_seed(seed)
    vm, Ks = S.make_rig(1, W, H)
    image_infos = {
        "img_idx": torch.full((H, W), img_idx, dtype=torch.long),
        "normed_time": torch.full((H, W), img_idx / (N_IMAGES - 1)),
        "pixels": torch.rand(H, W, 3, generator=g),
        "sky_masks": (torch.rand(H, W, generator=g) < 0.3).float(),
        "lidar_depth_map": torch.rand(H, W, generator=g) * 20.0 * (torch.rand(H, W, generator=g) < 0.2).float(),
        "egocar_masks": (torch.rand(H, W, generator=g) < 0.05).float(),
    }
    cam_infos = {
        "camera_to_world": torch.linalg.inv(vm[0]),
        "intrinsics": Ks[0],
        "height": torch.tensor(H, dtype=torch.long),
        "width": torch.tensor(W, dtype=torch.long),
    }
    to = lambda d: {k: v.to(device) for k, v in d.items()}  # noqa: E731
    return to(image_infos), to(cam_infos)


def all_params(trainer):
    out = {}
    for mname, m in trainer.models.items():
        for pname, p in m.named_parameters():
            out[f"{mname}.{pname}"] = p
    return out


# ---- CPU stand-ins for the kernel-backed entry points (the CPU test checks the trainer's HOST logic) ---------------
def oracle_render_fused(params, viewmats, Ks, width, height, sky=None, grid_slots=None, bil_sizes=(), sh_degree=3,
                        near_plane=0.1, far_plane=1e10, radius_clip=0.0, absgrad=True, row_begin=0, row_end=-1,
                        activated=False, dense_info=False, antialiased=False, guidance_factor=None):
    """``render.render_fused(activated=True)`` restated with the CPU oracle."""
    from oracle import bilateral_ref as B
    from oracle import raster_ref as R

    assert activated and viewmats.shape[0] == 1
    renders, alphas, info = R.rasterization(
        params["_means"], params["_quats"], params["_scales"], params["_opacities"].reshape(-1), params["_rgbs"],
        viewmats, Ks, width, height, near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip,
        render_mode="RGB+ED", rasterize_mode="antialiased" if antialiased else "classic", absgrad=absgrad)
    rgb_g = torch.clamp(renders[0, ..., :3], max=1.0)
    depth, alpha = renders[0, ..., 3:4], alphas[0]
    rgb_in = rgb_g if sky is None else rgb_g + sky * (1.0 - alpha)
    rgb = rgb_in if grid_slots is None else B.multiscale_forward(list(grid_slots[0]), rgb_in, guidance_factor)
    m2d = info["means2d"]
    if absgrad:
        m2d.absgrad = info["absgrad"]
    return dict(rgb=rgb, rgb_gaussians=rgb_g, depth=depth, opacity=alpha, radii=info["radii"].to(torch.int32),
                means2d=m2d, info=dict(n_isect=sum(info["n_isect"]), cache=None), pixel_rows=(0, height))


def oracle_multiscale_bilateral(rgb, slots, sizes, factors=(4, 4, 2), return_affine=False):
    from oracle import bilateral_ref as B

    affs = B.multiscale_affines(list(slots), rgb, factors)
    out = B.apply_chain(rgb, affs)
    if return_affine:
        H, W = rgb.shape[:2]
        return out, [a.reshape(1, H, W, 3, 4) for a in affs]
    return out


def oracle_tv(x, weight=1.0):
    from oracle import bilateral_ref as B

    return weight * B.total_variation_loss(x)


def oracle_tv_levels(grids, weights):
    return sum(oracle_tv(g, w) for g, w in zip(grids, weights))
