"""Stand-ins for the third-party packages the reference's hot-path files import but that are not installed in this
image (no network).  TEST INFRASTRUCTURE ONLY - our own code, nothing copied.

None of them carries arithmetic of the path under test, except two that are given small honest implementations so
that the reference's ``compute_losses`` runs end to end:

* ``omegaconf.OmegaConf``    attribute-style nested dict with the handful of methods the trainer uses
                             (``create / get / copy / update / pop / items``)
* ``pytorch_msssim.SSIM``    a plain 11x11 Gaussian-window SSIM (same definition, no multi-scale)

Everything else (tensorly, pytorch3d, nvdiffrast, kornia, viser, nerfview, torchmetrics, open3d, the dataset
package) is an empty shell whose functions raise when called.
"""
import sys
import types

import torch
from torch import nn


def _stub(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__bds_stub__", False):
        return sys.modules[name]          # the real package is installed: leave it alone
    m = sys.modules.get(name) or types.ModuleType(name)
    m.__bds_stub__ = True
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _not_available(*_a, **_k):  # pragma: no cover
    raise RuntimeError("stubbed third-party function called; it is not on the path under test")


# ---- omegaconf -------------------------------------------------------------------------------------------------
class Cfg(dict):
    """Attribute-style config node (the subset of omegaconf.DictConfig the reference trainer touches)."""

    def __init__(self, data=None):
        super().__init__()
        for k, v in (data or {}).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, Cfg):
            return Cfg(v)
        if isinstance(v, (list, tuple)):
            return [Cfg._wrap(x) for x in v]
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, Cfg._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        import copy

        return Cfg(copy.deepcopy(dict(self)))

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            self[k] = v

    def items(self):
        return list(super().items())     # a snapshot: MultiTrainer._init_models pops and re-inserts while iterating


class OmegaConf:
    @staticmethod
    def create(data=None):
        return Cfg(data or {})

    @staticmethod
    def to_container(cfg, **_k):
        return dict(cfg)


# ---- pytorch_msssim ----------------------------------------------------------------------------------------------
class SSIM(nn.Module):
    """Single-scale SSIM, 11x11 Gaussian window (sigma 1.5), K = (0.01, 0.03), valid convolution, mean over the map."""

    def __init__(self, data_range=1.0, size_average=True, channel=3, win_size=11, win_sigma=1.5):
        super().__init__()
        x = torch.arange(win_size, dtype=torch.float32) - win_size // 2
        g = torch.exp(-(x ** 2) / (2 * win_sigma ** 2))
        g = g / g.sum()
        self.register_buffer("win", (g[:, None] * g[None, :])[None, None].repeat(channel, 1, 1, 1))
        self.channel, self.data_range = channel, data_range

    def forward(self, X, Y):
        import torch.nn.functional as F

        win = self.win.to(X.dtype)
        C1, C2 = (0.01 * self.data_range) ** 2, (0.03 * self.data_range) ** 2
        f = lambda t: F.conv2d(t, win, groups=self.channel)  # noqa: E731
        mx, my = f(X), f(Y)
        sxx, syy, sxy = f(X * X) - mx * mx, f(Y * Y) - my * my, f(X * Y) - mx * my
        ssim = ((2 * mx * my + C1) * (2 * sxy + C2)) / ((mx * mx + my * my + C1) * (sxx + syy + C2))
        return ssim.mean()


class _PSNR(nn.Module):
    def __init__(self, data_range=1.0, **_k):
        super().__init__()
        self.data_range = data_range

    def forward(self, a, b):
        return 10.0 * torch.log10(self.data_range ** 2 / torch.mean((a - b) ** 2))


class _LPIPS(nn.Module):
    """NOT LPIPS (the network weights are not available offline): a deterministic stand-in (mean absolute difference)
    so that the reference's eval harness (models/video_utils.py:271-275) can be driven end to end."""

    def __init__(self, **_k):
        super().__init__()

    def forward(self, a, b):
        return (a - b).abs().mean()


# ---- skimage.metrics.structural_similarity (models/video_utils.py:11, 264-345) -----------------------------------------
def structural_similarity(im1, im2, *, data_range=None, channel_axis=None, full=False, win_size=7, **_k):
    """scikit-image's default SSIM restated (Wang et al. 2004 as implemented by skimage 0.2x): 7x7 uniform window,
    K1 = 0.01, K2 = 0.03, sample covariance (N / (N - 1)), mean over the map cropped by (win_size - 1) // 2 on each
    side, channels averaged.  ``full=True`` also returns the per-pixel map ([H, W, C] for multichannel input)."""
    import numpy as np
    from scipy.ndimage import uniform_filter

    if channel_axis is not None:
        im1, im2 = np.moveaxis(im1, channel_axis, -1), np.moveaxis(im2, channel_axis, -1)
        res = [structural_similarity(im1[..., c], im2[..., c], data_range=data_range, full=True, win_size=win_size)
               for c in range(im1.shape[-1])]
        mssim = float(np.mean([r[0] for r in res]))
        return (mssim, np.stack([r[1] for r in res], -1)) if full else mssim
    x, y = im1.astype(np.float64), im2.astype(np.float64)
    NP = win_size ** x.ndim
    cov_norm = NP / (NP - 1)
    f = lambda a: uniform_filter(a, size=win_size)  # noqa: E731
    ux, uy = f(x), f(y)
    vx, vy, vxy = cov_norm * (f(x * x) - ux * ux), cov_norm * (f(y * y) - uy * uy), cov_norm * (f(x * y) - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    mssim = float(S[tuple(slice(pad, -pad) for _ in range(x.ndim))].mean())
    return (mssim, S) if full else mssim


class _SplitWrapper:  # type annotation only in models/video_utils.py
    pass


class _DrivingDataset:  # type annotation only in models/trainers/scene_graph.py
    pass


def install():
    """Puts every stand-in into ``sys.modules`` (packages that ARE installed are left alone)."""
    _stub("tensorly", set_backend=lambda *_a, **_k: None)
    _stub("tensorly.decomposition", parafac=_not_available)
    p3d = _stub("pytorch3d")
    p3d.ops = _stub("pytorch3d.ops", knn_points=_not_available)
    p3d.transforms = _stub("pytorch3d.transforms", matrix_to_quaternion=_not_available)
    nvd = _stub("nvdiffrast")
    nvd.torch = _stub("nvdiffrast.torch", texture=_not_available)
    _stub("omegaconf", OmegaConf=OmegaConf, DictConfig=Cfg)
    _stub("kornia", losses=types.SimpleNamespace(inverse_depth_smoothness_loss=_not_available))
    _stub("viser", ViserServer=_not_available)
    _stub("nerfview", CameraState=object, Viewer=_not_available)   # annotation / viewer only
    _stub("open3d")
    _stub("pytorch_msssim", SSIM=SSIM)
    tm = _stub("torchmetrics")
    tm.image = _stub("torchmetrics.image", PeakSignalNoiseRatio=_PSNR)
    tm.image.lpip = _stub("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_LPIPS)
    ds = _stub("datasets")
    ds.__path__ = []   # a package, so that "datasets.driving_dataset" resolves to the stub below
    ds.driving_dataset = _stub("datasets.driving_dataset", DrivingDataset=_DrivingDataset)
    ds.base = _stub("datasets.base", SplitWrapper=_SplitWrapper)
    # the eval harness (models/video_utils.py) and utils/visualization.py: video writing / colour maps are not exercised
    _stub("imageio", mimwrite=_not_available, get_writer=_not_available)
    _stub("cv2")
    mpl = _stub("matplotlib")
    mpl.cm = _stub("matplotlib.cm", get_cmap=_not_available)
    sk = _stub("skimage")
    sk.metrics = _stub("skimage.metrics", structural_similarity=structural_similarity)
