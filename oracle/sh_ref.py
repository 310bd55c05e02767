"""CPU restatement of gsplat v1.3.0 ``spherical_harmonics`` (real SH, Sloan's fast evaluation).
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: gsplat is a pip dependency of the reference
(README.md:81, call sites models/gaussians/vanilla.py:383-389, nodes/rigid.py:462, ...), its source
is not under /root/reference and it is not installable here.  Constants and band layout are the
published ones.

coeffs [..., K, 3]; dirs [..., 3] (normalised inside, as the gsplat kernel does); bands above
``degree`` are ignored.  Returns [..., 3] WITHOUT the +0.5 / clamp, which the reference applies in
Python (vanilla.py:389).
"""
import torch


def sh_bases(degree: int, dirs):
    """Returns [..., (degree+1)^2] basis values for normalised dirs."""
    d = dirs / dirs.norm(dim=-1, keepdim=True)
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    out = [torch.full_like(x, 0.2820947917738781)]
    if degree >= 1:
        c1 = 0.48860251190292
        out += [-c1 * y, c1 * z, -c1 * x]
    if degree >= 2:
        z2 = z * z
        fTmp0B = -1.092548430592079 * z
        fC1 = x * x - y * y
        fS1 = 2 * x * y
        out += [
            0.5462742152960395 * fS1,
            fTmp0B * y,
            0.9461746957575601 * z2 - 0.3153915652525201,
            fTmp0B * x,
            0.5462742152960395 * fC1,
        ]
    if degree >= 3:
        fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658
        fTmp1B = 1.445305721320277 * z
        fC2 = x * fC1 - y * fS1
        fS2 = x * fS1 + y * fC1
        out += [
            -0.5900435899266435 * fS2,
            fTmp1B * fS1,
            fTmp0C * y,
            z * (1.865881662950577 * z2 - 1.119528997770346),
            fTmp0C * x,
            fTmp1B * fC1,
            -0.5900435899266435 * fC2,
        ]
    if degree >= 4:
        raise NotImplementedError("reference uses sh_degree <= 3 (configs/omnire_ms_bilateral.yaml:57)")
    return torch.stack(out, dim=-1)


def spherical_harmonics(degree: int, dirs, coeffs):
    b = sh_bases(degree, dirs)  # [..., nb]
    nb = b.shape[-1]
    return (b[..., :, None] * coeffs[..., :nb, :]).sum(-2)


def num_sh_bases(degree: int) -> int:
    return (degree + 1) ** 2
