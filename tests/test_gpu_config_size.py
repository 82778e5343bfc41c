"""Parity at the sizes BASELINE.json quotes ("What's weak" 1): the tensor-core path that produces the headline number
(z-march conv at 64^3: 144 CTAs, nine cost-balanced z-segments) against the CPU oracle itself, not only against the library's own
fp32 mode; the batch >= 4 kernel sequence of config 4 (grouped statistics off); a config-3-shaped volume (64^3 patches, stride 32)
with a ragged two-rank shard; and the FiLM-table reallocation regression.

The oracle forward of the driver U-Net at 64^3 takes ~1.5 s on the GPU box's host cores, so everything here stays well under a minute."""
import numpy as np
import pytest
import torch

from cases import MIN_BOUND, _unet
from diffusioniqt_b200 import volume as V
from diffusioniqt_b200.synth import synthetic_field, synthetic_noise, synthetic_state_dict
from helpers import max_rel, rel_err, spec_from_kwargs
from oracle import stitch_oracle as so
from oracle.ddpm_oracle import ddpm_sample
from oracle.unet_oracle import unet_forward

pytestmark = pytest.mark.gpu

DRIVER = _unet(64)          # train.py:83-116 + config/config.yaml (dim 64, mults 1-2-4, 2 resnet blocks / level, SE, deep_feature off)
CONFIGS = {"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}


def _driver_unet(seed, size=64):
    from diffusioniqt_b200 import Unet
    unet = Unet(**DRIVER, img_size=size)
    sd = synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=seed)
    unet.load_state_dict(sd)
    return unet, sd


def _imagen(unet, T, size=64, norm="z-score"):
    from diffusioniqt_b200 import Imagen, NullUnet
    return Imagen(unets=(NullUnet(), unet), configs={"Data": {"norm": norm}, "Train": {"batch_sample": False}}, image_sizes=(size, size),
                  channels=1, min_bound=MIN_BOUND, timesteps=T, pred_objectives="x_start", dynamic_thresholding=False, cond_drop_prob=0.0,
                  auto_normalize_img=False).cuda()


def test_config2_forward_bf16_against_the_oracle_with_taps():
    """BASELINE config 2 shape: the driver U-Net on one 64^3 patch, bf16 tcgen05 path, against oracle.unet_forward (fp32 CPU)."""
    from diffusioniqt_b200 import lib
    unet, sd = _driver_unet(seed=11)
    unet = unet.cuda().set_compute_dtype("bf16")
    unet.debug_taps = ("downs.0.1", "ups.1.1")
    x, lr = synthetic_field((1, 1, 64, 64, 64), 3), synthetic_field((1, 1, 64, 64, 64), 4)
    t = torch.tensor([1.3])
    got = unet(x.cuda(), None, t.cuda(), lowres_cond_img=lr.cuda()).cpu()
    eng = next(iter(unet._engines.values()))
    # the Block.project convs of the 64^3 and 32^3 levels (32 of the 38) run the z-march kernel, those of the 16^3 level the per-tap kernel
    # with split-K, everything else with C % 64 == 0 the per-tap kernel, init_conv its own fused tcgen05 kernel
    zm = [k for k, v in eng.conv_impls.items() if v == lib.IMPL_ZM]
    assert len(zm) == 32 and all(k.endswith(".project") for k in zm)
    assert len(eng.split_k) == 6 and eng.init_conv_fused
    # GroupNorm + FiLM + Mish ride on the conv's load path at the 64^3 and 32^3 levels
    assert len(eng.fused_gn) == 20 + 12, len(eng.fused_gn)
    taps = {}
    with torch.no_grad():
        want = unet_forward(sd, spec_from_kwargs(DRIVER), x, t, lowres_cond_img=lr, taps=taps)
    errs = {name: rel_err(eng.tap(name).cpu(), taps[name]) for name in unet.debug_taps}
    errs["final_conv"] = rel_err(got, want)
    print("config-2 forward, bf16 vs oracle, rel-L2:", {k: f"{v:.3e}" for k, v in errs.items()})
    # one ResnetBlock (2 convs, 2 GroupNorms, SE) deep: the north star's per-kernel bf16 tolerance; ups.1.1 sits behind ~30 bf16 kernels
    assert errs["downs.0.1"] < 1e-2
    assert errs["ups.1.1"] < 2e-2
    assert errs["final_conv"] < 3e-2
    assert max_rel(eng.tap("downs.0.1").cpu(), taps["downs.0.1"]) < 2e-2


def test_config2_sampler_t8_bf16_against_the_oracle():
    """Eight denoising iterations at config-2 size (the captured graph, the fused final conv + DDPM update) against the oracle sampler."""
    unet, sd = _driver_unet(seed=12)
    T, shape = 8, (1, 1, 64, 64, 64)
    imagen = _imagen(unet, T)
    imagen.unets[1].set_compute_dtype("bf16")
    lr = synthetic_field(shape, 5)
    noise = synthetic_noise(shape, T + 1, 6)
    imagen.noise_override = noise
    got, _, _ = imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)
    spec = spec_from_kwargs(DRIVER)
    with torch.no_grad():
        want, _, _ = ddpm_sample(lambda x, ls: unet_forward(sd, spec, x, ls, lowres_cond_img=lr), shape, noise, timesteps=T, min_bound=MIN_BOUND)
    err = rel_err(got.cpu(), want)
    print(f"config-2 sampler T=8, bf16 vs oracle, rel-L2: {err:.3e}")
    assert err < 6e-2                                     # the full-sampler bf16 tolerance of tests/test_gpu_sampler.py
    assert float(got.min()) >= MIN_BOUND - 1e-6


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-3), ("bf16", 6e-2)])
def test_config4_shape_elucidated_batch4_ungrouped(dtype, tol):
    """BASELINE config 4 runs batch 32: with batch > 2 the engine switches the grouped statistics OFF (separate finalize / SE-gate
    kernels), a different launch sequence from the benchmarked B = 1.  Batch 4 at 16^3 is what the CPU oracle can follow."""
    from cases import elucidated_hparams, elucidated_noise_count
    from diffusioniqt_b200 import ElucidatedImagen, NullUnet
    from helpers import oracle_elucidated
    case = dict(unet=DRIVER, batch=4, size=16, weight_seed=75, input_seed=85, noise_seed=95, hp=dict(num_sample_steps=4, sigma_max=20.0, S_churn=40.0),
                dynamic_threshold=False)
    unet, sd = _driver_unet(seed=case["weight_seed"], size=16)
    hp = elucidated_hparams(case)
    im = ElucidatedImagen(unets=(NullUnet(), unet), image_sizes=(16, 16), channels=1, cond_drop_prob=0.0, auto_normalize_img=False,
                          dynamic_thresholding=False, **hp).cuda()
    im.unets[1].set_compute_dtype(dtype)
    im.noise_override = synthetic_noise((4, 1, 16, 16, 16), elucidated_noise_count(case), case["noise_seed"])
    from cases import build_inputs
    _, lr, _ = build_inputs(case)
    img = im.sample(batch_size=4, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)
    eng = next(iter(im.unets[1]._engines.values()))
    assert eng.n == 4 and eng.grouped is False
    want, _ = oracle_elucidated(case, sd=sd)
    err = max_rel(img.cpu(), want) if dtype == "fp32" else rel_err(img.cpu(), want)
    print(f"config-4 shape (batch 4, ungrouped statistics) {dtype}: {err:.3e}")
    assert err < tol


def test_config3_shape_volume_ragged_two_rank_shard():
    """Config-3 geometry (64^3 patches, stride 32, crop margin 16) on a 128^3 volume with an air slab, fp32 mode, T = 2:
      * the per-patch sampler outputs match the oracle sampler on the same patches and noise;
      * the stitched volume equals the loop-by-loop restatement of test_all.py:239-300 applied to those patches, bit for bit;
      * a two-rank shard with an odd number of kept patches (one rank one short, padded) stitches to the same bits."""
    P, stride, T, N = 64, 32, 2, 128
    unet, sd = _driver_unet(seed=13)
    imagen = _imagen(unet, T)
    imagen.unets[1].set_compute_dtype("fp32")
    low = synthetic_field((N, N, N), 21)
    low[:100] = low.min()                                          # air: only the nine patches starting at i = 64 pass the 5 % rule
    raw = low - low.min()
    grid = V.patch_grid(low.shape, P, stride)
    kept = [g for g in grid if not so.is_skipped(raw.numpy(), list(g), P)]
    assert 0 < len(kept) < len(grid) and len(kept) % 2 == 1, (len(kept), len(grid))
    noise = {g: synthetic_noise((1, 1, P, P, P), T + 1, 300 + n) for n, g in enumerate(kept)}
    outs = {}

    def run(rank, world, gather_fn=None):
        start, stop, _ = V.shard_range(len(kept), rank, world)
        order = iter(kept[start:stop])

        def sample_fn(lr):
            g = next(order)
            imagen.noise_override = noise[g]
            out = imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0]
            outs[g] = out[0, 0].cpu()
            return out

        return V.infer_volume(sample_fn, low.cuda(), patch=P, overlap=stride, raw_lowres=raw.cuda(), batch_size=1, fill_value=MIN_BOUND,
                              rank=rank, world=world, gather_fn=gather_fn)

    single = run(0, 1)
    assert single.n_patches == len(kept) and single.n_skipped == len(grid) - len(kept)
    got = single.volume.cpu()

    # (1) per-patch parity with the oracle sampler on three of the kept patches (first, middle, last)
    spec = spec_from_kwargs(DRIVER)
    for g in (kept[0], kept[len(kept) // 2], kept[-1]):
        lr = low[g[0]:g[0] + P, g[1]:g[1] + P, g[2]:g[2] + P][None, None]
        with torch.no_grad():
            want, _, _ = ddpm_sample(lambda x, ls: unet_forward(sd, spec, x, ls, lowres_cond_img=lr), (1, 1, P, P, P), noise[g], timesteps=T,
                                     min_bound=MIN_BOUND)
        assert max_rel(outs[g], want[0, 0]) < 2e-3, g

    # (2) stitch: the device kernel against the loop-by-loop restatement on the same patches
    want_vol = np.full((N, N, N), MIN_BOUND, np.float32)
    so.stitch(want_vol, [outs[g].numpy() for g in kept], [list(g) for g in kept], P, stride, False)
    want_vol = torch.from_numpy(so.background_mask(want_vol, low.numpy()))
    assert torch.equal(got, want_vol)

    # (3) two ranks, emulated in one process: each rank's padded shard is collected, then rank 0 stitches the concatenation
    shards = {}

    def collect(rank):
        def fn(local):
            shards[rank] = local.clone()
            return None                                            # infer_volume stops before the stitch
        return fn

    for r in (0, 1):
        assert run(r, 2, gather_fn=collect(r)) is None
    per = V.shard_range(len(kept), 0, 2)[2]
    assert shards[0].shape[0] == shards[1].shape[0] == per and 2 * per == len(kept) + 1      # ragged: the last rank is one short
    two = V.infer_volume(lambda lr: torch.zeros_like(lr), low.cuda(), patch=P, overlap=stride, raw_lowres=raw.cuda(), batch_size=1,
                         fill_value=MIN_BOUND, rank=0, world=2, gather_fn=lambda local: torch.cat([shards[0], shards[1]]))
    assert torch.equal(two.volume.cpu(), got)


def test_film_table_reallocation_drops_stale_graphs():
    """a later sampler with more steps reallocates the shared FiLM table; graphs captured on the old table must be
    re-captured instead of replayed on freed memory."""
    from cases import SAMPLE_CASES, build_inputs
    from helpers import weights_for
    from diffusioniqt_b200 import Imagen, NullUnet, Unet
    case = SAMPLE_CASES["skip_dim32_s8_t20_skip4"]
    unet = Unet(**dict(case["unet"], img_size=case["size"]))
    unet.load_state_dict(weights_for(case))
    im = Imagen(unets=(NullUnet(), unet), configs=CONFIGS, image_sizes=(8, 8), channels=1, min_bound=MIN_BOUND, timesteps=20,
                pred_objectives="x_start", dynamic_thresholding=False, cond_drop_prob=0.0, auto_normalize_img=False).cuda()
    _, lr, _ = build_inputs(case)
    shape = (1, 1, 8, 8, 8)
    short, full = synthetic_noise(shape, 7, 56), synthetic_noise(shape, 21, 57)

    def go(skip, nz):
        im.noise_override = nz
        return im.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, skip_steps=skip, use_tqdm=False)[0].clone()

    a = go(4, short)                    # 6 FiLM rows
    eng = next(iter(im.unets[1]._engines.values()))
    gen0 = eng.film_gen
    b = go(None, full)                  # 20 rows: the table is reallocated
    assert eng.film_gen == gen0 + 1
    filler = [torch.randn(1 << 20, device="cuda") for _ in range(8)]   # let the allocator hand the freed block to someone else
    a2 = go(4, short)
    b2 = go(None, full)
    del filler
    assert torch.equal(a, a2) and torch.equal(b, b2)


def test_config2_size_training_gradients_bf16_against_autograd_of_the_oracle():
    """SURVEY 8 f-4 at the BASELINE config-2 shape: one training evaluation of the driver U-Net on a 64^3 patch in bf16 (tcgen05 forward /
    data-gradient convs, tcgen05 weight gradients with 144-CTA voxel chunking, the reverse passes over 32 MiB tensors) against PyTorch
    autograd over the fp32 CPU oracle.  The gradient as one vector over all parameters, and the largest tensors one by one."""
    import torch.nn.functional as F
    from diffusioniqt_b200 import lib as L
    from diffusioniqt_b200.train import UnetBackprop
    unet, sd = _driver_unet(seed=11)
    unet = unet.cuda().set_compute_dtype("bf16")
    x, lr = synthetic_field((1, 1, 64, 64, 64), 3), synthetic_field((1, 1, 64, 64, 64), 4)
    t = torch.tensor([1.3])
    target = synthetic_field((1, 1, 64, 64, 64), 5)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    out = unet_forward(sdg, spec_from_kwargs(DRIVER), x, t, lowres_cond_img=lr)
    F.mse_loss(out, target).backward()
    bp = UnetBackprop(unet)
    n0 = L.launch_count()
    pred = bp.forward(x.cuda(), t.cuda(), lowres_cond_img=lr.cuda())
    assert rel_err(pred.cpu(), out.detach()) < 3e-2
    bp.backward(2 * (pred - target.cuda()) / pred.numel())
    assert L.launch_count() - n0 > 500                       # the step ran on this library's kernels
    params = dict(unet.named_parameters())
    names = [k for k, v in sdg.items() if v.requires_grad and v.grad is not None]
    got = torch.cat([params[k].grad.cpu().double().reshape(-1) for k in names])
    want = torch.cat([sdg[k].grad.double().reshape(-1) for k in names])
    assert rel_err(got, want) < 8e-2
    big = sorted(names, key=lambda k: -sdg[k].grad.norm().item())[:12]
    for k in big:
        assert rel_err(params[k].grad.cpu(), sdg[k].grad) < 8e-2, k
