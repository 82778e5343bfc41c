"""CPU oracle for the acceptance metrics  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates /root/reference/metrics.py:17-30 as used by test_all.py:47-85: PSNR after min-max normalisation of both
volumes (data_range 1), and SSIM with a 3-D gaussian window (torchmetrics 0.9.0 `StructuralSimilarityIndexMeasure`
defaults: kernel 11, sigma 1.5, k1 0.01, k2 0.03).  torchmetrics is a third-party dependency that is absent here
(requirements.txt:201 pins 0.9.0), so its published formula is restated and the SAME restatement is applied to both
sides of every comparison; absolute values are not claimed to match torchmetrics bit for bit (parity unpinned).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _minmax(x, rng=None):
    lo, hi = (x.min(), x.max()) if rng is None else rng
    return (x - lo) / (hi - lo)


def psnr(pred: torch.Tensor, target: torch.Tensor, rng=None) -> float:
    """metrics.py:17-21.  rng=None: each volume is scaled by its OWN min / max, as the reference does.  rng=(lo, hi): both
    volumes are scaled by the same fixed range (used by tests that must not hinge on one extreme voxel)."""
    p, t = _minmax(pred.double(), rng), _minmax(target.double(), rng)
    mse = torch.mean((p - t) ** 2)
    return float(10.0 * torch.log10(1.0 / mse))


def ssim3d(pred: torch.Tensor, target: torch.Tensor, kernel_size: int = 3, sigma: float = 1.5, normalise: bool = True, rng=None) -> float:
    p, t = pred.double(), target.double()
    if normalise:
        p, t = _minmax(p, rng), _minmax(t, rng)
    p, t = p[None, None], t[None, None]
    g = torch.arange(kernel_size, dtype=torch.float64) - (kernel_size - 1) / 2
    g = torch.exp(-(g ** 2) / (2 * sigma ** 2))
    g = g / g.sum()
    k3 = (g[:, None, None] * g[None, :, None] * g[None, None, :])[None, None]
    c1, c2 = 0.01 ** 2, 0.03 ** 2

    def blur(x):
        return F.conv3d(x, k3)

    mu_p, mu_t = blur(p), blur(t)
    s_pp = blur(p * p) - mu_p ** 2
    s_tt = blur(t * t) - mu_t ** 2
    s_pt = blur(p * t) - mu_p * mu_t
    ssim = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (s_pp + s_tt + c2))
    return float(ssim.mean())
