"""Elucidated (Karras / Heun) sampler on the GPU against the oracle and the reference-loop fixtures."""
import pytest
import torch

from cases import ELUCIDATED_CASES, build_inputs, elucidated_hparams, elucidated_noise_count
from diffusioniqt_b200.synth import synthetic_noise
from helpers import load_golden, max_rel, oracle_elucidated, rel_err, weights_for

pytestmark = pytest.mark.gpu


def _sampler(case, dtype):
    from diffusioniqt_b200 import ElucidatedImagen, NullUnet, Unet
    unet = Unet(**dict(case["unet"], img_size=case["size"]))
    unet.load_state_dict(weights_for(case))
    S = case["size"]
    hp = elucidated_hparams(case)
    im = ElucidatedImagen(unets=(NullUnet(), unet), image_sizes=(S, S), channels=1, cond_drop_prob=0.0, auto_normalize_img=False,
                          dynamic_thresholding=case["dynamic_threshold"], **hp).cuda()
    im.unets[1].set_compute_dtype(dtype)
    return im


def _noise(case):
    B, S = case["batch"], case["size"]
    return synthetic_noise((B, 1, S, S, S), elucidated_noise_count(case), case["noise_seed"])


def _run(im, case, **kw):
    _, lr, _ = build_inputs(case)
    return im.sample(batch_size=case["batch"], start_image_or_video=lr, start_at_unet_number=2, skip_steps=case.get("skip_steps"),
                     use_tqdm=False, **kw)


@pytest.mark.parametrize("name", list(ELUCIDATED_CASES))
def test_elucidated_fp32_matches_oracle_and_reference_fixture(name):
    case = ELUCIDATED_CASES[name]
    im = _sampler(case, "fp32")
    im.noise_override = _noise(case)
    img = _run(im, case)
    want, x_starts = oracle_elucidated(case)
    # fp32 mode, CUDA-core accumulation order: same tolerance as the DDPM sampler (tests/test_gpu_sampler.py)
    assert max_rel(img.cpu(), want) < 2e-3
    assert max_rel(img.cpu(), load_golden(name)["img"]) < 2e-3
    assert max_rel(im.last_x_start.cpu(), x_starts[-1]) < 2e-3
    assert float(img.min()) >= -1.0 and float(img.max()) <= 1.0


@pytest.mark.parametrize("name", ["edm_dim32_s8_n6", "edm_driver_dim64_s8_n4_b2"])
def test_elucidated_bf16_close_to_oracle(name):
    case = ELUCIDATED_CASES[name]
    im = _sampler(case, "bf16")
    im.noise_override = _noise(case)
    img = _run(im, case)
    want, _ = oracle_elucidated(case)
    assert rel_err(img.cpu(), want) < 6e-2


def test_elucidated_graph_replay_equals_eager_launches():
    case = ELUCIDATED_CASES["edm_dim32_s8_n6"]
    outs = []
    for use_graph in (True, False):
        im = _sampler(case, "bf16")
        im.use_cuda_graph = use_graph
        im.noise_override = _noise(case)
        outs.append(_run(im, case))
    assert torch.equal(outs[0], outs[1])


def test_elucidated_rng_draw_order_and_reuse():
    """randn(shape) once, then one randn per step (elucidated_imagen.py:432, 476); a second call reuses the captured graphs."""
    case = ELUCIDATED_CASES["edm_dim32_s8_n6"]
    B, S = case["batch"], case["size"]
    im = _sampler(case, "bf16")
    torch.manual_seed(99)
    a = _run(im, case)
    torch.manual_seed(99)
    im.noise_override = [torch.randn((B, 1, S, S, S), device="cuda") for _ in range(elucidated_noise_count(case))]
    b = _run(im, case)
    assert torch.equal(a, b)
    im.noise_override = None
    torch.manual_seed(99)
    assert torch.equal(a, _run(im, case))


def test_elucidated_sigma_overrides_and_clamp_range():
    case = ELUCIDATED_CASES["edm_dim32_s8_n6"]
    im = _sampler(case, "fp32")
    im.clamp_range = (-0.5, 2.0)
    im.noise_override = _noise(case)
    img = _run(im, case, sigma_max=10.0)
    from oracle.elucidated_oracle import elucidated_sample
    from oracle.unet_oracle import unet_forward
    from helpers import spec_from_kwargs
    sd, spec = weights_for(case), spec_from_kwargs(case["unet"])
    _, lr, _ = build_inputs(case)
    hp = {k: v for k, v in elucidated_hparams(case).items() if k not in ("P_mean", "P_std")}
    hp["sigma_max"] = 10.0
    with torch.no_grad():
        want, _ = elucidated_sample(lambda x, t: unet_forward(sd, spec, x, t, lowres_cond_img=lr), (1, 1, 8, 8, 8), _noise(case),
                                    dynamic_threshold=False, clamp_range=(-0.5, 2.0), **hp)
    assert max_rel(img.cpu(), want) < 2e-3
    assert float(img.min()) >= -0.5
