"""CPU oracle for the DDPM sampler around the U-Net  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, in plain PyTorch fp32, the continuous-time cosine schedule and the ancestral sampling
loop of the reference: `GaussianDiffusionContinuousTimes` (/root/reference/imagen_pytorch3D.py:
229-234, 236-309) and `Imagen.p_sample_loop / p_sample / p_mean_variance` (:2058-2160, :2032-2056,
:1976-2030) for `pred_objective` in {'x_start', 'noise', 'v'}, static clamping (z-score:
`clamp_(min=min_bound)`, min-max: `clamp_(-1, 1)`) and optional dynamic thresholding.

Noise is injected: `noise[0]` is the initial `randn(shape)` (:2080) and `noise[1 + i]` the
`randn_like` of step i (:2051, drawn on every step including the last), so that the reference,
this oracle and the CUDA sampler can be driven by one recorded sequence.

Pinned against the live reference / committed fixtures by tests/test_oracle_vs_reference.py and
tests/test_oracle_golden.py (the reference has no golden vectors of its own).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


def alpha_cosine_log_snr(t: Tensor, s: float = 0.008) -> Tensor:
    # imagen_pytorch3D.py:229-231 ; log(x, eps) = log(clamp(x, min=eps))
    return -torch.log(((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1).clamp(min=1e-5))


def beta_linear_log_snr(t: Tensor) -> Tensor:
    # imagen_pytorch3D.py:225-227
    return -torch.log(torch.special.expm1(1e-4 + 10 * (t ** 2)))


def log_snr_to_alpha_sigma(log_snr: Tensor) -> Tuple[Tensor, Tensor]:
    # imagen_pytorch3D.py:233-234
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


def sampling_timesteps(num_timesteps: int, skip_steps: Optional[int] = None) -> List[Tuple[float, float]]:
    """(t, t_next) pairs of get_sampling_timesteps (:261-266) after the skip rule (:2103-2107)."""
    times = torch.linspace(1.0, 0.0, num_timesteps + 1)
    pairs = list(zip(times[:-1].tolist(), times[1:].tolist()))
    skip = skip_steps or 0
    if skip > 1:
        pairs = pairs[::skip] + [pairs[-1]]
    return pairs


def q_posterior(x_start: Tensor, x_t: Tensor, t: Tensor, t_next: Tensor, log_snr_fn=alpha_cosine_log_snr):
    # imagen_pytorch3D.py:290-309
    shape = (-1,) + (1,) * (x_t.dim() - 1)
    log_snr = log_snr_fn(t).reshape(shape)
    log_snr_next = log_snr_fn(t_next).reshape(shape)
    alpha, sigma = log_snr_to_alpha_sigma(log_snr)
    alpha_next, sigma_next = log_snr_to_alpha_sigma(log_snr_next)
    c = -torch.special.expm1(log_snr - log_snr_next)
    mean = alpha_next * (x_t * (1 - c) / alpha + c * x_start)
    var = (sigma_next ** 2) * c
    log_var = torch.log(var.clamp(min=1e-20))
    return mean, var, log_var


def ddpm_sample(unet_fn: Callable[[Tensor, Tensor], Tensor], shape: Sequence[int], noise: Sequence[Tensor], *,
                timesteps: int, min_bound: float, norm: str = "z-score", pred_objective: str = "x_start",
                dynamic_threshold: bool = False, dynamic_thresholding_percentile: float = 0.95,
                skip_steps: Optional[int] = None, init_images: Optional[Tensor] = None,
                noise_schedule: str = "cosine"):
    """Returns (img, [x_t after each step] + [final], [x_start of each step] + [last]).

    `unet_fn(x_t, log_snr)` is the network (already bound to its low-res conditioning)."""
    log_snr_fn = alpha_cosine_log_snr if noise_schedule == "cosine" else beta_linear_log_snr
    b = shape[0]
    img = noise[0].clone()                                                   # :2080
    if init_images is not None:
        img = img + init_images                                             # :2084-2085
    traj_x, traj_x0 = [], []
    x_start = None
    for i, (tv, tnv) in enumerate(sampling_timesteps(timesteps, skip_steps)):
        t = torch.full((b,), tv, dtype=torch.float32)
        t_next = torch.full((b,), tnv, dtype=torch.float32)
        pred = unet_fn(img, log_snr_fn(t))                                   # :1994
        pad = (-1,) + (1,) * (img.dim() - 1)
        if pred_objective == "x_start":
            x_start = pred
        elif pred_objective == "noise":                                      # :354-357
            alpha, sigma = log_snr_to_alpha_sigma(log_snr_fn(t).reshape(pad))
            x_start = (img - sigma * pred) / alpha.clamp(min=1e-8)
        elif pred_objective == "v":                                          # :347-351
            alpha, sigma = log_snr_to_alpha_sigma(log_snr_fn(t).reshape(pad))
            x_start = alpha * img - sigma * pred
        else:
            raise ValueError(f"unknown objective {pred_objective}")
        if dynamic_threshold:                                                # :2006-2021
            s = torch.quantile(x_start.reshape(b, -1).abs(), dynamic_thresholding_percentile, dim=-1)
            s = s.clamp(min=1.0) if norm == "min-max" else s.clamp(min=min_bound)
            s = s.reshape(pad)
            x_start = x_start.clamp(-s, s) / s
        elif norm == "min-max":
            x_start = x_start.clamp(-1.0, 1.0)                               # :2024
        else:
            x_start = x_start.clamp(min=min_bound)                           # :2026
        mean, _, log_var = q_posterior(x_start, img, t, t_next, log_snr_fn)  # :2029
        nonzero = (1 - (t_next == 0).float()).reshape(pad)                   # :2053-2054
        img = mean + nonzero * (0.5 * log_var).exp() * noise[1 + i]          # :2055
        traj_x.append(img.clone())
        traj_x0.append(x_start.clone())
    traj_x.append(img.clone())                                               # :2151-2152
    traj_x0.append(x_start.clone())
    img = img.clamp(-1.0, 1.0) if norm == "min-max" else img.clamp(min=min_bound)  # :2154-2157
    return img, traj_x, traj_x0
