"""Case table shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
from __future__ import annotations

import torch

from diffusioniqt_b200.synth import synthetic_field

MIN_BOUND = (0.0 - 271.64814106698583) / 377.117173547721      # train.py:72 with config.yaml:12-13


def _unet(dim, **kw):
    base = dict(dim=dim, init_dim=dim, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1,
                lowres_cond=True, init_cross_embed=False, attend_at_middle=False,
                attend_at_enc=(False, False, False), use_se_attn=True, pixel_shuffle_upsample=True,
                memory_efficient=False, deep_feature=False, boundary=False, batch_sample=False)
    base.update(kw)
    return base


def _attn(att_type, **kw):
    base = dict(att_type=att_type, attend_at_enc=(True, True, True), attend_at_enc_depth=(1, 2, 1), attend_at_enc_heads=(2, 4, 2),
                attend_at_middle=True, attend_at_middle_depth=1, attend_at_middle_heads=2, attn_dim_head=16, att_forward_expansion=2,
                att_drop=0.0, att_forward_drop=0.0, att_localvit=False, deep_feature=True, batch_sample=True, num_resnet_blocks=(1, 1, 1),
                img_size=24)
    base.update(kw)
    return base


# U-Net forward cases: {name: {unet kwargs, batch, size, seeds, log-SNR values, hooked submodules}}
FORWARD_CASES = {
    # driver architecture (train.py:83-116) at a CPU-sized patch
    "driver_dim64_s16": dict(unet=_unet(64), batch=1, size=16, weight_seed=11, input_seed=21,
                             log_snr=[2.5], taps=("init_conv", "downs.0.1", "downs.0.4", "downs.2.4", "ups.0.0", "ups.0.1", "final_res_block")),
    # BASELINE config 1 architecture (dim 32), two different noise levels in one batch
    "cfg1_dim32_s16_b2": dict(unet=_unet(32), batch=2, size=16, weight_seed=12, input_seed=22,
                              log_snr=[-4.0, 6.0], taps=("downs.1.3.1", "ups.1.2.1")),
    # deep_feature=True executes the mid block (imagen_pytorch3D.py:1633-1651)
    "deep_dim32_s8": dict(unet=_unet(32, deep_feature=True), batch=1, size=8, weight_seed=13, input_seed=23,
                          log_snr=[0.3], taps=("mid_block",)),
    # eval_config.yaml geometry: 27 sub-volumes with boundary exchange (imagen_pytorch3D.py:37-46)
    "boundary_dim32_s8": dict(unet=_unet(32, boundary=True, batch_sample=True), batch=27, size=8, weight_seed=14,
                              input_seed=24, log_snr=[1.0] * 27, taps=("downs.0.1",)),
    # different depth / width pattern and skip scaling
    "alt_dim32_s16": dict(unet=_unet(32, dim_mults=(1, 2), num_resnet_blocks=(1, 2), scale_skip_connection=True, init_dim=64),
                          batch=1, size=16, weight_seed=15, input_seed=25, log_snr=[-1.0], taps=()),
    # ConvTranspose3d upsampling (Upsample_deconv :440-457, pixel_shuffle_upsample=False)
    "deconv_dim32_s16": dict(unet=_unet(32, pixel_shuffle_upsample=False), batch=2, size=16, weight_seed=23, input_seed=33, log_snr=[0.7, -1.2],
                             taps=("ups.0.0", "ups.1.0")),
    # the same with channel counts that are multiples of 64 (tcgen05 convs in bf16 mode)
    "deconv_dim64_s16": dict(unet=_unet(64, pixel_shuffle_upsample=False), batch=1, size=16, weight_seed=24, input_seed=34, log_snr=[0.2], taps=("ups.0.0",)),
    # the constructor-default init conv: CrossEmbedLayer with kernel sizes (3, 7, 15) (imagen_pytorch3D.py:661-686, 1222-1223)
    "crossembed_dim32_s16": dict(unet=_unet(32, init_cross_embed=True, init_cross_embed_kernel_sizes=(3, 7, 15)), batch=1, size=16, weight_seed=22,
                                 input_seed=32, log_snr=[0.4], taps=("init_conv",)),
    # ---- attention blocks (off in the shipped configs).  27 sub-volumes of 8^3 = one merged 24^3 volume; patch sizes
    # 8 / 4 / 2 / 2 give 27 tokens at every level.  `img_size` is the merged side (it only sizes the ViT position table).
    "attn_linear_dim32_s8": dict(unet=_unet(32, **_attn("linear")), batch=27, size=8, weight_seed=16, input_seed=26,
                                 log_snr=[0.7] * 27, taps=("downs.0.2", "downs.1.2", "mid_attn")),
    "attn_softmax_boundary_dim32_s8": dict(unet=_unet(32, **_attn("softmax", boundary=True)), batch=27, size=8, weight_seed=17, input_seed=27,
                                           log_snr=[-0.5] * 27, taps=("downs.0.2", "mid_attn")),
    "attn_vit_dim32_s8": dict(unet=_unet(32, **_attn("vit", att_localvit=False)), batch=27, size=8, weight_seed=18, input_seed=28,
                              log_snr=[1.5] * 27, taps=("downs.0.2", "mid_attn")),
    "attn_vitlocal_dim32_s8": dict(unet=_unet(32, **_attn("vit", att_localvit=True, attend_at_enc=(True, False, True))), batch=27, size=8,
                                   weight_seed=19, input_seed=29, log_snr=[0.0] * 27, taps=("downs.2.2",)),
    # channel counts that are multiples of 64 (tcgen05 1x1x1 convs in bf16), 2^3 sub-volumes of 32^3 -> 512 tokens, head dim 64
    "attn_linear_dim64_f2_s32": dict(unet=_unet(64, **_attn("linear", batch_sample_factor=2, img_size=64, attn_dim_head=64, attend_at_enc_depth=(1, 1, 1),
                                                            attend_at_enc_heads=(2, 2, 4), attend_at_enc=(True, False, True))),
                                     batch=8, size=32, weight_seed=20, input_seed=30, log_snr=[0.9] * 8, taps=("downs.0.2",)),
    # softmax attention with head dim 64: the tcgen05 attention kernel (csrc/attn_tc.cu) in bf16 mode
    "attn_softmax_dim64_f2_s32": dict(unet=_unet(64, **_attn("softmax", batch_sample_factor=2, img_size=64, attn_dim_head=64, attend_at_enc_depth=(1, 1, 1),
                                                             attend_at_enc_heads=(2, 2, 2), attend_at_enc=(True, True, False), attend_at_middle_heads=4)),
                                      batch=8, size=32, weight_seed=21, input_seed=31, log_snr=[-0.3] * 8, taps=("mid_attn",)),
}

# Full-sampler cases (Imagen.sample with injected noise)
SAMPLE_CASES = {
    "cfg1_dim32_s16_t12": dict(unet=_unet(32), batch=1, size=16, timesteps=12, weight_seed=31, input_seed=41,
                               noise_seed=51, min_bound=MIN_BOUND, norm="z-score"),
    "driver_dim64_s8_t6_b2": dict(unet=_unet(64), batch=2, size=8, timesteps=6, weight_seed=32, input_seed=42,
                                  noise_seed=52, min_bound=MIN_BOUND, norm="z-score"),
    "minmax_dim32_s8_t8": dict(unet=_unet(32), batch=1, size=8, timesteps=8, weight_seed=33, input_seed=43,
                               noise_seed=53, min_bound=MIN_BOUND, norm="min-max"),
    "noise_obj_dim32_s8_t8": dict(unet=_unet(32), batch=1, size=8, timesteps=8, weight_seed=34, input_seed=44,
                                  noise_seed=54, min_bound=MIN_BOUND, norm="min-max", pred_objective="noise"),
    "boundary_dim32_s8_t4": dict(unet=_unet(32, boundary=True, batch_sample=True), batch=27, size=8, timesteps=4,
                                 weight_seed=35, input_seed=45, noise_seed=55, min_bound=MIN_BOUND, norm="z-score",
                                 boundary=True),
    # BASELINE.json configs[0] exactly: the reference's own CPU-runnable case (dim 32, one 32^3 patch, 50-step DDPM, fp32)
    "baseline_cfg1_dim32_s32_t50": dict(unet=_unet(32), batch=1, size=32, timesteps=50, weight_seed=37, input_seed=47,
                                        noise_seed=57, min_bound=MIN_BOUND, norm="z-score"),
    "skip_dim32_s8_t20_skip4": dict(unet=_unet(32), batch=1, size=8, timesteps=20, skip_steps=4, weight_seed=36,
                                    input_seed=46, noise_seed=56, min_bound=MIN_BOUND, norm="z-score"),
}


def unet_kwargs_for_reference(case):
    kw = dict(case["unet"])
    kw.setdefault("img_size", case["size"])
    return kw


def make_configs(case):
    """The nested dict the reference reads inside the sampler (imagen_pytorch3D.py:2016, 2023, 2154)."""
    return {"Data": {"norm": case.get("norm", "z-score")}, "Train": {"batch_sample": case["unet"].get("batch_sample", False)}}


def build_inputs(case):
    B, S = case["batch"], case["size"]
    x = synthetic_field((B, 1, S, S, S), case["input_seed"])
    lr = synthetic_field((B, 1, S, S, S), case["input_seed"] + 1000)
    time = torch.tensor(case.get("log_snr", [0.0] * B), dtype=torch.float32)
    return x, lr, time


def sample_noise_count(case):
    T = case["timesteps"]
    skip = case.get("skip_steps")
    return (len(range(0, T, skip)) + 1 if skip and skip > 1 else T) + 1


def tap_digest(v):
    """Strided sub-sample of an activation (B, C, D, H, W): every 8th channel, every 2nd voxel."""
    return v[:1, ::8, ::2, ::2, ::2].contiguous()


# Elucidated (Karras / Heun) sampler cases: reference `ElucidatedImagen.one_unet_sample` around the 3-D U-Net (adapter of
# SURVEY.md Appendix C).  Fewer steps / milder sigma range than the defaults keep the CPU fixtures small; `hp` overrides
# elucidated_imagen.py:96-106.
ELUCIDATED_CASES = {
    "edm_dim32_s8_n6": dict(unet=_unet(32), batch=1, size=8, weight_seed=71, input_seed=81, noise_seed=91,
                            hp=dict(num_sample_steps=6), dynamic_threshold=False),
    "edm_driver_dim64_s8_n4_b2": dict(unet=_unet(64), batch=2, size=8, weight_seed=72, input_seed=82, noise_seed=92,
                                      hp=dict(num_sample_steps=4, sigma_max=20.0, S_churn=40.0), dynamic_threshold=False),
    "edm_dynthr_dim32_s8_n5": dict(unet=_unet(32), batch=1, size=8, weight_seed=73, input_seed=83, noise_seed=93,
                                   hp=dict(num_sample_steps=5), dynamic_threshold=True),
    "edm_skip_dim32_s8_n8_skip3": dict(unet=_unet(32), batch=1, size=8, weight_seed=74, input_seed=84, noise_seed=94,
                                       hp=dict(num_sample_steps=8), dynamic_threshold=False, skip_steps=3),
}


def elucidated_hparams(case):
    hp = dict(num_sample_steps=32, sigma_min=0.002, sigma_max=80.0, sigma_data=0.5, rho=7.0, P_mean=-1.2, P_std=1.2, S_churn=80.0,
              S_tmin=0.05, S_tmax=50.0, S_noise=1.003)
    hp.update(case.get("hp", {}))
    return hp


def elucidated_noise_count(case):
    return elucidated_hparams(case)["num_sample_steps"] - (case.get("skip_steps") or 0) + 1
