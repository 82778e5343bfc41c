"""Quality metrics of the reference's inference script: metrics.py:17-34 as driven by test_all.py:47-85.

PSNR after min-max normalisation of both volumes (data_range 1), SSIM and multi-scale SSIM with a 3-D gaussian window.  The window
metrics come from torchmetrics (pinned 0.9.0, requirements.txt:201), a third-party dependency that is not vendored in the reference and
is absent here; its published formulas are restated (separable gaussian 11 / sigma 1.5, k1 0.01, k2 0.03, valid region, MS-SSIM betas
0.0448, 0.2856, 0.3001, 0.2363, 0.1333 with 2x average pooling between scales).  Parity with torchmetrics is unpinned.  LPIPS (a
pretrained VGG) is out of scope.  Plain torch tensor ops: this is post-processing around the hot path, it runs on whatever device
holds the volumes.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

MS_SSIM_BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _as5d(x) -> torch.Tensor:
    x = torch.as_tensor(x)
    while x.dim() < 5:
        x = x[None]
    return x.double()


def _minmax(x: torch.Tensor) -> torch.Tensor:
    return (x - x.min()) / (x.max() - x.min())


def psnr(pred, target) -> float:
    """metrics.py:17-21: both volumes scaled by their own min / max, data_range 1."""
    p, t = _minmax(_as5d(pred)), _minmax(_as5d(target))
    return float(10.0 * torch.log10(1.0 / torch.mean((p - t) ** 2)))


def _gauss(kernel_size: int, sigma: float, device) -> torch.Tensor:
    g = torch.arange(kernel_size, dtype=torch.float64, device=device) - (kernel_size - 1) / 2
    g = torch.exp(-(g ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _blur(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """Separable 3-D gaussian over the valid region."""
    k = g.numel()
    x = F.conv3d(x, g.reshape(1, 1, k, 1, 1))
    x = F.conv3d(x, g.reshape(1, 1, 1, k, 1))
    return F.conv3d(x, g.reshape(1, 1, 1, 1, k))


def _ssim_and_cs(p: torch.Tensor, t: torch.Tensor, kernel_size: int, sigma: float, data_range: float) -> Tuple[torch.Tensor, torch.Tensor]:
    g = _gauss(kernel_size, sigma, p.device)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu_p, mu_t = _blur(p, g), _blur(t, g)
    s_pp = _blur(p * p, g) - mu_p ** 2
    s_tt = _blur(t * t, g) - mu_t ** 2
    s_pt = _blur(p * t, g) - mu_p * mu_t
    cs = (2 * s_pt + c2) / (s_pp + s_tt + c2)
    ssim = ((2 * mu_p * mu_t + c1) / (mu_p ** 2 + mu_t ** 2 + c1)) * cs
    return ssim.mean(), cs.mean()


def ssim3d(pred, target, kernel_size: int = 3, sigma: float = 1.5, data_range: Optional[float] = None) -> float:
    """metrics.py:23-30 (the reference's SSIM() defaults to a 3-wide window).  data_range=None: min-max normalise both volumes first (as the reference does) and use range 1."""
    p, t = _as5d(pred), _as5d(target)
    if data_range is None:
        p, t, data_range = _minmax(p), _minmax(t), 1.0
    return float(_ssim_and_cs(p, t, kernel_size, sigma, data_range)[0])


def ms_ssim3d(pred, target, kernel_size: int = 11, sigma: float = 1.5, data_range: float = 1.0, betas: Sequence[float] = MS_SSIM_BETAS,
              normalize: Optional[str] = None) -> float:
    """metrics.py:32-34 `MultiScaleStructuralSimilarityIndexMeasure()`: prod_i cs_i^beta_i (all but the last scale) * ssim_last^beta_last."""
    p, t = _as5d(pred), _as5d(target)
    smallest = min(p.shape[-3:])
    if smallest // 2 ** (len(betas) - 1) < kernel_size:
        raise ValueError(f"volume side {smallest} is too small for {len(betas)} scales with an {kernel_size}-wide window "
                         f"(needs at least {kernel_size * 2 ** (len(betas) - 1)})")
    sims, css = [], []
    for _ in betas:
        s, c = _ssim_and_cs(p, t, kernel_size, sigma, data_range)
        if normalize == "relu":
            s, c = torch.relu(s), torch.relu(c)
        sims.append(s)
        css.append(c)
        p, t = F.avg_pool3d(p, 2), F.avg_pool3d(t, 2)
    sims, css = torch.stack(sims), torch.stack(css)
    if normalize == "simple":
        sims, css = (sims + 1) / 2, (css + 1) / 2
    b = torch.tensor(betas, dtype=torch.float64, device=sims.device)
    return float(torch.prod(css[:-1] ** b[:-1]) * sims[-1] ** b[-1])


def evaluate(gt, pred, kernel_size: int = 11) -> Tuple[float, float]:
    """test_all.py:47-85 without LPIPS: crop 24 (240-voxel volumes) or 32 (256) voxels per face, then (MS-SSIM of the min-max
    normalised volumes, PSNR)."""
    gt, pred = torch.as_tensor(gt), torch.as_tensor(pred)
    m = {240: 24, 256: 32}.get(gt.shape[0], 0)
    if m:
        gt, pred = gt[m:-m, m:-m, m:-m], pred[m:-m, m:-m, m:-m]
    p = psnr(gt, pred)
    s = ms_ssim3d(_minmax(_as5d(gt)), _minmax(_as5d(pred)), kernel_size=kernel_size)
    return s, p
