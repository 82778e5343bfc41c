"""CPU oracle for the Elucidated (Karras et al. / Heun) sampler  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, in plain PyTorch fp32, `ElucidatedImagen.one_unet_sample` (/root/reference/elucidated_imagen.py:382-532)
with `sample_schedule` (:365-379), the preconditioning coefficients (:308-324), `preconditioned_network_forward`
(:329-358) and `threshold_x_start` (:298-311).

The reference class cannot be CONSTRUCTED with the 3-D `Unet` of imagen_pytorch3D.py, nor call it (SURVEY.md
Appendix C: `cast_model_parameters` rejects `cond_on_text`; the U-Net takes `(x, time_steps, time)` and no
`text_embeds` / `lowres_noise_times`; sigma is padded for 4-D images).  The adapter this oracle (and the product) uses
is the minimal one of Appendix C: 5-D padding of sigma, the call `unet(c_in * x, <unused>, c_noise(sigma),
lowres_cond_img=lr)`, and no noising of the low-res conditioning (the U-Net was trained on the clean low-field patch,
imagen_pytorch3D.py:2303-2304).

Pinning: tests/golden/make_golden_elucidated.py runs the reference's OWN, unmodified `one_unet_sample` /
`preconditioned_network_forward` / `sample_schedule` methods on an instance assembled without the broken constructor,
around the reference `Unet`, with exactly that adapter; tests/test_oracle_golden.py compares this file with the
committed fixture and tests/test_oracle_vs_reference.py with the live reference when /root/reference exists.

Noise is injected: `noise[0]` is the initial `randn(shape)` (:432) and `noise[1 + i]` the `randn` of step i (:476).
"""
from __future__ import annotations

from math import sqrt
from typing import Callable, Optional, Sequence

import torch

Tensor = torch.Tensor

DEFAULTS = dict(num_sample_steps=32, sigma_min=0.002, sigma_max=80.0, sigma_data=0.5, rho=7.0, S_churn=80.0, S_tmin=0.05,
                S_tmax=50.0, S_noise=1.003)      # elucidated_imagen.py:96-106


def _log(t: Tensor, eps: float = 1e-20) -> Tensor:
    return torch.log(t.clamp(min=eps))           # :67-68


def sample_schedule(num_sample_steps: int, rho: float, sigma_min: float, sigma_max: float) -> Tensor:
    # :365-379
    N = num_sample_steps
    inv_rho = 1 / rho
    steps = torch.arange(num_sample_steps, dtype=torch.float32)
    sigmas = (sigma_max ** inv_rho + steps / (N - 1) * (sigma_min ** inv_rho - sigma_max ** inv_rho)) ** rho
    return torch.nn.functional.pad(sigmas, (0, 1), value=0.)


def c_skip(sigma_data, sigma):
    return (sigma_data ** 2) / (sigma ** 2 + sigma_data ** 2)                    # :310-311


def c_out(sigma_data, sigma):
    return sigma * sigma_data * (sigma_data ** 2 + sigma ** 2) ** -0.5           # :313-314


def c_in(sigma_data, sigma):
    return 1 * (sigma ** 2 + sigma_data ** 2) ** -0.5                            # :316-317


def c_noise(sigma):
    return _log(sigma) * 0.25                                                    # :319-320


def threshold_x_start(x_start: Tensor, dynamic_threshold: bool, percentile: float = 0.95, clamp_range=(-1.0, 1.0)) -> Tensor:
    # :298-311 (clamp_range generalises the literal clamp(-1, 1) for z-score data)
    if not dynamic_threshold:
        return x_start.clamp(clamp_range[0], clamp_range[1])
    s = torch.quantile(x_start.reshape(x_start.shape[0], -1).abs(), percentile, dim=-1)
    s = s.clamp(min=1.)
    s = s.reshape(-1, *((1,) * (x_start.dim() - 1)))
    return x_start.clamp(-s, s) / s


def preconditioned_forward(unet_fn: Callable[[Tensor, Tensor], Tensor], x: Tensor, sigma: float, *, sigma_data: float, clamp: bool,
                           dynamic_threshold: bool, percentile: float = 0.95, clamp_range=(-1.0, 1.0)) -> Tensor:
    # :329-358, with sigma padded to 5-D
    b = x.shape[0]
    sig = torch.full((b,), sigma, dtype=torch.float32)
    pad = sig.reshape(b, *((1,) * (x.dim() - 1)))
    net_out = unet_fn(c_in(sigma_data, pad) * x, c_noise(sig))
    out = c_skip(sigma_data, pad) * x + c_out(sigma_data, pad) * net_out
    if not clamp:
        return out
    return threshold_x_start(out, dynamic_threshold, percentile, clamp_range)


def elucidated_sample(unet_fn: Callable[[Tensor, Tensor], Tensor], shape: Sequence[int], noise: Sequence[Tensor], *,
                      num_sample_steps: int = 32, sigma_min: float = 0.002, sigma_max: float = 80.0, sigma_data: float = 0.5,
                      rho: float = 7.0, S_churn: float = 80.0, S_tmin: float = 0.05, S_tmax: float = 50.0, S_noise: float = 1.003,
                      clamp: bool = True, dynamic_threshold: bool = False, percentile: float = 0.95,
                      init_images: Optional[Tensor] = None, skip_steps: Optional[int] = None, clamp_range=(-1.0, 1.0)):
    """`unet_fn(x_scaled, c_noise)` is the network already bound to its low-res conditioning.
    Returns (images, [x_start estimate after each step])."""
    sigmas = sample_schedule(num_sample_steps, rho, sigma_min, sigma_max)
    gammas = torch.where((sigmas >= S_tmin) & (sigmas <= S_tmax), min(S_churn / num_sample_steps, sqrt(2) - 1), 0.)   # :418-422
    sched = list(zip(sigmas[:-1], sigmas[1:], gammas[:-1]))
    images = sigmas[0] * noise[0]                                                # :430-432
    if init_images is not None:
        images = images + init_images                                            # :436-437
    sched = sched[(skip_steps or 0):]                                            # :463-464
    kw = dict(sigma_data=sigma_data, clamp=clamp, dynamic_threshold=dynamic_threshold, percentile=percentile, clamp_range=clamp_range)
    x_starts = []
    for ind, (sigma, sigma_next, gamma) in enumerate(sched):
        sigma, sigma_next, gamma = (t.item() for t in (sigma, sigma_next, gamma))  # :471
        eps = S_noise * noise[1 + ind]                                           # :476
        sigma_hat = sigma + gamma * sigma
        added_noise = sqrt(sigma_hat ** 2 - sigma ** 2) * eps
        images_hat = images + added_noise                                        # :478-481
        model_output = preconditioned_forward(unet_fn, images_hat, sigma_hat, **kw)             # :488-494
        denoised_over_sigma = (images_hat - model_output) / sigma_hat            # :496
        images_next = images_hat + (sigma_next - sigma_hat) * denoised_over_sigma  # :498
        x_start = model_output
        if sigma_next != 0:                                                      # :502-516
            model_output_next = preconditioned_forward(unet_fn, images_next, sigma_next, **kw)
            denoised_prime_over_sigma = (images_next - model_output_next) / sigma_next
            images_next = images_hat + 0.5 * (sigma_next - sigma_hat) * (denoised_over_sigma + denoised_prime_over_sigma)
            x_start = model_output_next
        images = images_next
        x_starts.append(x_start.clone())
    images = images.clamp(clamp_range[0], clamp_range[1])                        # :527
    return images, x_starts
