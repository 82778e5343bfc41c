"""Whole-volume path on the GPU: patches -> Imagen.sample (CUDA kernels) -> stitch, against the CPU oracle pipeline,
judged with the north star's acceptance metrics (PSNR within 0.05 dB, SSIM within 1e-3 of the reference pipeline)."""
import numpy as np
import pytest
import torch

from cases import MIN_BOUND
from diffusioniqt_b200 import volume as V
from diffusioniqt_b200.synth import synthetic_field, synthetic_noise, synthetic_state_dict
from helpers import spec_from_kwargs
from oracle import metrics_oracle as mo
from oracle import stitch_oracle as so
from oracle.ddpm_oracle import ddpm_sample
from oracle.unet_oracle import unet_forward

pytestmark = pytest.mark.gpu

KW = dict(dim=32, init_dim=32, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, init_cross_embed=False,
          attend_at_middle=False, attend_at_enc=(False, False, False), use_se_attn=True, memory_efficient=False, deep_feature=False,
          boundary=False, batch_sample=False)


@pytest.mark.parametrize("dtype,norm,psnr_tol,ssim_tol", [("fp32", "z-score", 0.05, 1e-3), ("bf16", "z-score", 0.05, 1e-3), ("bf16", "min-max", 0.05, 1e-3)])
def test_stitched_volume_psnr_ssim(dtype, norm, psnr_tol, ssim_tol):
    from diffusioniqt_b200 import Imagen, NullUnet, Unet
    P, stride, T, N = 16, 8, 6, 32
    unet = Unet(**KW, img_size=P)
    sd = synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=61)
    unet.load_state_dict(sd)
    imagen = Imagen(unets=(NullUnet(), unet), configs={"Data": {"norm": norm}, "Train": {"batch_sample": False}}, image_sizes=(P, P),
                    channels=1, min_bound=MIN_BOUND, timesteps=T, pred_objectives="x_start", dynamic_thresholding=False,
                    cond_drop_prob=0.0).cuda()
    imagen.unets[1].set_compute_dtype(dtype)
    lowres = synthetic_field((N, N, N), 62)[...]
    lowres[:6, :10] = lowres.min()                       # a background corner
    grid = V.patch_grid(lowres.shape, P, stride)
    noise = {g: synthetic_noise((1, 1, P, P, P), T + 1, 64 + n) for n, g in enumerate(grid)}
    order = iter(grid)

    def gpu_sampler(lr):
        imagen.noise_override = noise[next(order)]
        return imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0]

    res = V.infer_volume(gpu_sampler, lowres.cuda(), patch=P, overlap=stride, raw_lowres=(lowres - lowres.min()).cuda(), batch_size=1,
                         fill_value=MIN_BOUND)
    got = res.volume.cpu()

    spec = spec_from_kwargs(KW)
    outs, kept = [], []
    raw = (lowres - lowres.min()).numpy()
    for g in grid:
        if so.is_skipped(raw, list(g), P):
            continue
        lr = lowres[g[0]:g[0] + P, g[1]:g[1] + P, g[2]:g[2] + P][None, None]
        with torch.no_grad():
            img, _, _ = ddpm_sample(lambda x, ls: unet_forward(sd, spec, x, ls, lowres_cond_img=lr), (1, 1, P, P, P), noise[g], timesteps=T,
                                    min_bound=MIN_BOUND, norm=norm)
        outs.append(img[0, 0].numpy())
        kept.append(list(g))
    want = np.full((N, N, N), MIN_BOUND, np.float32)
    so.stitch(want, outs, kept, P, stride, False)
    want = torch.from_numpy(so.background_mask(want, lowres.numpy()))
    assert res.n_patches == len(kept)
    # No trained checkpoint ships, so the "ground truth" is synthetic: the reference pipeline's own output plus an independent field
    # at -25 dB of its range, clipped to that range (CPU-checked: the reference pipeline then scores ~24 dB / SSIM 0.88, the regime
    # of a real IQT model; against an unrelated random truth, ~13 dB, the statistic mostly measures chance correlations).
    lo, hi = float(want.min()), float(want.max())
    truth = (want + synthetic_field((N, N, N), 63) * ((hi - lo) * 10 ** (-25 / 20))).clamp(lo, hi)
    assert 20.0 < mo.psnr(want, truth) < 28.0
    # North-star acceptance: PSNR within 0.05 dB and SSIM within 1e-3 of the reference pipeline.  The reference's metric scales every
    # volume by its OWN min / max (metrics.py:18-19); with random weights the maximum is a single outlier voxel (8.1 against an rms of
    # 1.8), so in bf16 a 1 % change of that ONE voxel rescales the whole volume and moves the literal metric by tenths of a dB.  The
    # tolerance is therefore asserted with both volumes scaled by the truth's range, and the literal metric with a looser bound in bf16.
    rng = (truth.min(), truth.max())
    assert abs(mo.psnr(got, truth, rng) - mo.psnr(want, truth, rng)) < psnr_tol
    assert abs(mo.ssim3d(got, truth, rng=rng) - mo.ssim3d(want, truth, rng=rng)) < ssim_tol
    # In min-max mode (Data.norm = "min-max": x0 and the final patch are clamped to [-1, 1], :2024, :2154) the extremes of both volumes are
    # the clamp bounds themselves, as they are for a trained model whose output range is set by the data: there the LITERAL metric is held
    # to the north-star tolerance in bf16 too.  In z-score mode with random weights the measured bf16 deviation is printed and recorded in
    # DESIGN.md section 4.
    d_psnr, d_ssim = abs(mo.psnr(got, truth) - mo.psnr(want, truth)), abs(mo.ssim3d(got, truth) - mo.ssim3d(want, truth))
    print(f"literal metric deviation [{dtype}, {norm}]: dPSNR = {d_psnr:.4f} dB, dSSIM = {d_ssim:.2e}; "
          f"range-pinned: dPSNR = {abs(mo.psnr(got, truth, rng) - mo.psnr(want, truth, rng)):.4f} dB")
    literal = (1.0, 1e-2) if (dtype == "bf16" and norm == "z-score") else (psnr_tol, ssim_tol)
    assert d_psnr < literal[0]
    assert d_ssim < literal[1]
    assert mo.psnr(got, want) > (60.0 if dtype == "fp32" else 30.0)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_back_to_back_samples_are_reproducible(dtype):
    """Samples queued back to back without host synchronisation (CUDA-graph replay, programmatic dependent launch inside the graph)
    must not interfere with each other.  Every reduction has a fixed order, so repeats must be bit-identical."""
    from diffusioniqt_b200 import Imagen, NullUnet, Unet
    P, T = 16, 6
    unet = Unet(**KW, img_size=P)
    unet.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=61))
    imagen = Imagen(unets=(NullUnet(), unet), configs={"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}, image_sizes=(P, P),
                    channels=1, min_bound=MIN_BOUND, timesteps=T, pred_objectives="x_start", dynamic_thresholding=False,
                    cond_drop_prob=0.0).cuda()
    imagen.unets[1].set_compute_dtype(dtype)
    lrs = [synthetic_field((1, 1, P, P, P), 70 + i).cuda() for i in range(3)]
    noises = [[t.cuda() for t in synthetic_noise((1, 1, P, P, P), T + 1, 80 + i)] for i in range(3)]
    outs = []
    for rep in range(8):                                   # 24 samples queued back to back, no host sync in between
        for lr, nz in zip(lrs, noises):
            imagen.noise_override = nz
            outs.append(imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0])
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        assert torch.isfinite(o).all(), i
        assert torch.equal(o, outs[i % 3]), f"sample {i} differs from its first run"


def test_engine_and_graph_survive_repeated_sample_calls():
    """One engine build + one graph capture serve every sample() call of a volume (regression: `.to(device)` inside sample() used to
    drop the engine each call while the sampler cache, keyed by id(engine), could pair a dead engine's graph with the new buffers)."""
    from diffusioniqt_b200 import Imagen, NullUnet, Unet
    P, T = 16, 3
    unet = Unet(**KW, img_size=P)
    unet.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=61))
    imagen = Imagen(unets=(NullUnet(), unet), configs={"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}, image_sizes=(P, P),
                    channels=1, min_bound=MIN_BOUND, timesteps=T, pred_objectives="x_start", dynamic_thresholding=False,
                    cond_drop_prob=0.0).cuda()
    lr = synthetic_field((1, 1, P, P, P), 70).cuda()
    imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)
    eng = next(iter(imagen.unets[1]._engines.values()))
    st = next(iter(eng.sampler_cache.values()))
    graph = st.graph
    junk = [torch.empty(1 << 20, device="cuda") for _ in range(8)]        # perturb the allocator between calls
    for _ in range(3):
        imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)
        junk.pop()
    assert list(imagen.unets[1]._engines.values()) == [eng]
    assert next(iter(eng.sampler_cache.values())) is st and st.graph is graph


def test_file_to_file_inference(tmp_path):
    """NIfTI in -> z-score -> device patch grid -> Imagen.sample -> device stitch -> NIfTI out + metrics (test_all.py:182-316)."""
    from diffusioniqt_b200 import Imagen, ImagenTrainer, NullUnet, Unet
    from diffusioniqt_b200.infer import infer_nifti
    from diffusioniqt_b200.nifti import load_nifti, save_nifti
    P, T, N = 16, 3, 48
    unet = Unet(**KW, img_size=P)
    unet.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=61))
    configs = {"Data": {"norm": "z-score", "mean": 271.648, "std": 377.117}, "Train": {"batch_sample": False, "patch_size_sub": P},
               "Eval": {"overlap": 8, "batch_size": 4}}
    imagen = Imagen(unets=(NullUnet(), unet), configs=configs, image_sizes=(P, P), channels=1, min_bound=MIN_BOUND, timesteps=T,
                    pred_objectives="x_start", dynamic_thresholding=False, cond_drop_prob=0.0).cuda()
    trainer = ImagenTrainer(configs, imagen=imagen, use_ema=True)
    raw = (synthetic_field((N, N, N), 90) * 300 + 400).clamp(min=0).numpy().astype(np.float32)
    raw[:20] = 0.0                                   # air: the patches inside this slab are skipped (data.py:192-196)
    affine = np.diag([1.5, 1.5, 1.5, 1.0])
    save_nifti(raw, affine, tmp_path / "lr.nii.gz")
    save_nifti(raw * 1.1, affine, tmp_path / "hr.nii.gz")
    torch.manual_seed(7)
    res = infer_nifti(trainer, configs, tmp_path / "lr.nii.gz", tmp_path / "pred.nii.gz", tmp_path / "hr.nii.gz", evaluate_kernel_size=3)
    data, aff, _ = load_nifti(tmp_path / "pred.nii.gz")
    assert data.shape == (N, N, N) and np.allclose(aff, affine)
    assert np.array_equal(data.astype(np.float32), res.prediction.numpy())
    assert res.n_patches > 0 and res.n_skipped > 0 and np.isfinite(res.prediction.numpy()).all()
    low = (torch.from_numpy(raw) - 271.648) / 377.117
    assert torch.equal(res.prediction[low == low.min()], torch.full_like(res.prediction[low == low.min()], float(low.min())))
    assert float(res.prediction.min()) >= min(MIN_BOUND, float(low.min())) - 1e-6      # sampler clamp (:2157) / background / fill value
    assert res.psnr is not None and 0.0 < res.ms_ssim <= 1.0


def test_literal_acceptance_metric_with_trained_weights():
    """The north star's LITERAL criterion (per-volume PSNR / SSIM with every volume scaled by its own min / max, metrics.py:17-30) in bf16 and
    z-score mode on weights that have been TRAINED, not drawn at random: a short run of this package's own training step (Imagen.forward,
    the hand-written reverse pass, the Adam kernel) on synthetic high-field / low-field pairs teaches the U-Net to return a smooth field
    close to its conditioning image, so the extremes of the stitched volume are set by the data, not by a single outlier voxel.
    The reference pipeline (CPU oracle, fp32) runs on the same trained weights, noise and patches."""
    import torch.nn.functional as F
    from diffusioniqt_b200 import Imagen, ImagenTrainer, NullUnet, Unet
    P, stride, T, N = 16, 8, 6, 32
    unet = Unet(**KW, img_size=P)
    unet.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=61))
    configs = {"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}
    imagen = Imagen(unets=(NullUnet(), unet), configs=configs, image_sizes=(P, P), channels=1, min_bound=MIN_BOUND, timesteps=T, pred_objectives="x_start",
                    loss_type="l2", p2_loss_weight_gamma=0.5, dynamic_thresholding=False, cond_drop_prob=0.0).cuda()
    imagen.unets[1].set_compute_dtype("fp32")
    trainer = ImagenTrainer(configs=configs, imagen=imagen, lr=3e-4, use_ema=False, gradient_accumulation_steps=1, verbose=False)
    trainer.train()
    hr_vol = synthetic_field((48, 48, 48), 91, smooth=3).clamp(min=MIN_BOUND)

    def degrade(v):      # low-field stand-in: blurred, noisier
        b = F.avg_pool3d(v[None, None], 3, stride=1, padding=1)[0, 0]
        return b + 0.15 * synthetic_field(tuple(v.shape), 92, smooth=0)

    lr_vol = degrade(hr_vol)
    rs = np.random.RandomState(5)
    torch.manual_seed(5)
    losses = []
    for step in range(160):
        idx = rs.randint(0, 48 - P + 1, size=(4, 3))
        hr = torch.stack([hr_vol[i:i + P, j:j + P, k:k + P] for i, j, k in idx])[:, None]
        lo = torch.stack([lr_vol[i:i + P, j:j + P, k:k + P] for i, j, k in idx])[:, None]
        losses.append(trainer(hr, lo, unet_number=2)[0])
    assert np.mean(losses[-20:]) < 0.5 * np.mean(losses[:5]), (losses[:5], losses[-20:])
    sd = {k: v.detach().cpu().clone() for k, v in imagen.unets[1].state_dict().items()}

    # ---- the volume pipeline on the trained weights: bf16 kernels against the fp32 oracle
    imagen.eval()
    imagen.unets[1].set_compute_dtype("bf16")
    test_hr = synthetic_field((N, N, N), 93, smooth=3).clamp(min=MIN_BOUND)
    lowres = degrade(test_hr)
    lowres[:6, :10] = lowres.min()
    grid = V.patch_grid(lowres.shape, P, stride)
    noise = {g: synthetic_noise((1, 1, P, P, P), T + 1, 164 + n) for n, g in enumerate(grid)}
    order = iter(grid)

    def gpu_sampler(lr):
        imagen.noise_override = noise[next(order)]
        return imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0]

    res = V.infer_volume(gpu_sampler, lowres.cuda(), patch=P, overlap=stride, raw_lowres=(lowres - lowres.min()).cuda(), batch_size=1, fill_value=MIN_BOUND)
    got = res.volume.cpu()
    spec = spec_from_kwargs(KW)
    outs, kept = [], []
    raw = (lowres - lowres.min()).numpy()
    for g in grid:
        if so.is_skipped(raw, list(g), P):
            continue
        lr = lowres[g[0]:g[0] + P, g[1]:g[1] + P, g[2]:g[2] + P][None, None]
        with torch.no_grad():
            img, _, _ = ddpm_sample(lambda x, ls: unet_forward(sd, spec, x, ls, lowres_cond_img=lr), (1, 1, P, P, P), noise[g], timesteps=T,
                                    min_bound=MIN_BOUND, norm="z-score")
        outs.append(img[0, 0].numpy())
        kept.append(list(g))
    want = np.full((N, N, N), MIN_BOUND, np.float32)
    so.stitch(want, outs, kept, P, stride, False)
    want = torch.from_numpy(so.background_mask(want, lowres.numpy()))
    truth = torch.from_numpy(so.background_mask(test_hr.numpy().copy(), lowres.numpy()))      # the real high-field volume this time
    p_ref, p_got = mo.psnr(want, truth), mo.psnr(got, truth)
    s_ref, s_got = mo.ssim3d(want, truth), mo.ssim3d(got, truth)
    print(f"trained weights, bf16 z-score: reference pipeline PSNR {p_ref:.3f} dB SSIM {s_ref:.4f}; this library {p_got:.3f} dB {s_got:.4f}; "
          f"loss {np.mean(losses[:5]):.3f} -> {np.mean(losses[-20:]):.3f}")
    assert abs(p_got - p_ref) < 0.05          # the north star's literal tolerance
    assert abs(s_got - s_ref) < 1e-3
    assert mo.psnr(got, want) > 30.0
