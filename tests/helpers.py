"""Shared test helpers: build oracle specs / weights / samplers from the golden case table."""
from __future__ import annotations

import os

import numpy as np
import torch

from cases import build_inputs, sample_noise_count
from diffusioniqt_b200.synth import synthetic_noise, synthetic_state_dict
from oracle.ddpm_oracle import ddpm_sample
from oracle.unet_oracle import UnetSpec, unet_forward

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_SPEC_KEYS = ("dim", "init_dim", "dim_mults", "num_resnet_blocks", "resnet_groups", "channels", "channels_out",
              "lowres_cond", "cond_images_channels", "self_cond", "learned_sinu_pos_emb_dim", "use_se_attn",
              "scale_skip_connection", "final_resnet_block", "deep_feature", "boundary", "batch_sample_factor",
              "init_conv_kernel_size", "att_type", "attend_at_enc", "attend_at_enc_depth", "attend_at_enc_heads", "attend_at_middle",
              "attend_at_middle_depth", "attend_at_middle_heads", "attn_dim_head", "att_localvit")


def spec_from_kwargs(kw) -> UnetSpec:
    return UnetSpec(**{k: kw[k] for k in _SPEC_KEYS if k in kw})


def state_dict_shapes(kw):
    """Parameter names/shapes of the reference `Unet` for constructor kwargs `kw`, derived from
    this repo's own module (whose state_dict contract is tested against the reference)."""
    from diffusioniqt_b200 import Unet
    return {k: tuple(v.shape) for k, v in Unet(**kw).state_dict().items()}


def weights_for(case):
    return synthetic_state_dict(state_dict_shapes(case["unet"]), case["weight_seed"])


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def oracle_forward(case, sd=None, taps=None):
    sd = sd if sd is not None else weights_for(case)
    x, lr, time = build_inputs(case)
    with torch.no_grad():
        return unet_forward(sd, spec_from_kwargs(case["unet"]), x, time, lowres_cond_img=lr, taps=taps)


def oracle_sample(case, sd=None):
    sd = sd if sd is not None else weights_for(case)
    spec = spec_from_kwargs(case["unet"])
    _, lr, _ = build_inputs(case)
    B, S = case["batch"], case["size"]
    noise = synthetic_noise((B, 1, S, S, S), sample_noise_count(case), case["noise_seed"])
    with torch.no_grad():
        return ddpm_sample(lambda x, log_snr: unet_forward(sd, spec, x, log_snr, lowres_cond_img=lr),
                           (B, 1, S, S, S), noise, timesteps=case["timesteps"], min_bound=case["min_bound"],
                           norm=case.get("norm", "z-score"), pred_objective=case.get("pred_objective", "x_start"),
                           dynamic_threshold=case.get("dynamic_threshold", False), skip_steps=case.get("skip_steps"))


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


def max_rel(a, b):
    """max |a-b| / max |b|  (the 'relative' used for per-kernel tolerances)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def oracle_elucidated(case, sd=None):
    """Elucidated sampler oracle on a case of cases.ELUCIDATED_CASES -> (img, [x_start per step])."""
    from cases import elucidated_hparams, elucidated_noise_count
    from oracle.elucidated_oracle import elucidated_sample
    sd = sd if sd is not None else weights_for(case)
    spec = spec_from_kwargs(case["unet"])
    _, lr, _ = build_inputs(case)
    B, S = case["batch"], case["size"]
    hp = {k: v for k, v in elucidated_hparams(case).items() if k not in ("P_mean", "P_std")}
    noise = synthetic_noise((B, 1, S, S, S), elucidated_noise_count(case), case["noise_seed"])
    with torch.no_grad():
        return elucidated_sample(lambda x, t: unet_forward(sd, spec, x, t, lowres_cond_img=lr), (B, 1, S, S, S), noise,
                                 dynamic_threshold=case["dynamic_threshold"], skip_steps=case.get("skip_steps"), **hp)
