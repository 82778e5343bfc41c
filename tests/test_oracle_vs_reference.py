"""Pin the oracle against the live, unmodified reference (only where /root/reference exists)."""
import pytest
import torch

import ref_shim
from cases import FORWARD_CASES, SAMPLE_CASES, build_inputs, make_configs, sample_noise_count, unet_kwargs_for_reference
from diffusioniqt_b200.synth import fill_module_, synthetic_noise
from helpers import oracle_elucidated, oracle_forward, oracle_sample, spec_from_kwargs
from oracle import unet_oracle as uo

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


@pytest.mark.parametrize("name", ["cfg1_dim32_s16_b2", "deep_dim32_s8", "boundary_dim32_s8", "alt_dim32_s16", "attn_linear_dim32_s8",
                                  "attn_softmax_boundary_dim32_s8", "attn_vit_dim32_s8", "attn_vitlocal_dim32_s8", "crossembed_dim32_s16", "deconv_dim32_s16"])
def test_unet_forward_bit_exact(ref, name):
    case = FORWARD_CASES[name]
    unet = ref.Unet(**unet_kwargs_for_reference(case)).eval()
    fill_module_(unet, seed=case["weight_seed"])
    x, lr, time = build_inputs(case)
    with torch.no_grad():
        want = unet(x, None, time, lowres_cond_img=lr)
    got = oracle_forward(case, sd=unet.state_dict())
    assert torch.equal(got, want)


def test_cross_embed_downsample_cannot_be_constructed_in_the_reference(ref):
    """Why diffusioniqt_b200.Unet refuses `cross_embed_downsample=True`: the reference passes dim_out into CrossEmbedLayer's `kernel_sizes`
    slot, which the partial at imagen_pytorch3D.py:1342 already fills."""
    kw = unet_kwargs_for_reference(FORWARD_CASES["cfg1_dim32_s16_b2"])
    with pytest.raises(TypeError, match="kernel_sizes"):
        ref.Unet(**{**kw, "cross_embed_downsample": True})


def test_sub_volume_helpers(ref):
    import sys
    sys.path.insert(0, ref_shim.REFERENCE_DIR)
    import utils_mine
    x = torch.randn(27, 3, 4, 4, 4)
    merged = utils_mine.merge_sub_volumes(x, (1, 3, 12, 12, 12))
    assert torch.equal(uo.merge_sub_volumes(x, 3), merged)
    assert torch.equal(uo.split_sub_volumes(merged, 3), utils_mine.convertVolume2subVolume(merged, (27, 3, 4, 4, 4)))
    assert torch.equal(uo.boundary_pad(x, 3), ref.boundary_pad(x))


def test_sampler_matches(ref):
    from make_golden import _InjectedNoise
    case = SAMPLE_CASES["minmax_dim32_s8_t8"]
    unet = ref.Unet(**unet_kwargs_for_reference(case)).eval()
    fill_module_(unet, seed=case["weight_seed"])
    S, B, T = case["size"], case["batch"], case["timesteps"]
    imagen = ref.Imagen(unets=(ref.NullUnet(), unet), configs=make_configs(case), image_sizes=(S, S), channels=1,
                        min_bound=case["min_bound"], timesteps=T, pred_objectives="x_start", dynamic_thresholding=False,
                        p2_loss_weight_gamma=0.0, auto_normalize_img=False, cond_drop_prob=0.0)
    _, lr, _ = build_inputs(case)
    noise = synthetic_noise((B, 1, S, S, S), sample_noise_count(case), case["noise_seed"])
    with _InjectedNoise(noise):
        want, _, traj = imagen.sample(batch_size=B, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)
    got, _, traj_o = oracle_sample(case, sd=unet.state_dict())
    assert torch.allclose(got, want, atol=1e-6, rtol=0)
    assert len(traj) == len(traj_o) == T + 1


def test_elucidated_sampler_matches_reference_loop(ref):
    """The reference's own one_unet_sample / preconditioned_network_forward (elucidated_imagen.py:329-532), run on an
    instance assembled without its broken constructor, against the oracle restatement: bit-identical on the same CPU."""
    import make_golden_elucidated as mg
    from cases import ELUCIDATED_CASES
    case = ELUCIDATED_CASES["edm_dim32_s8_n6"]
    want = mg.run_reference(case)
    got, _ = oracle_elucidated(case)
    assert torch.equal(got, want)


def test_flop_count_matches_survey():
    spec = uo.UnetSpec(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, deep_feature=False)
    assert abs(uo.count_flops(spec, 1, 64) / 1e9 - 1488.44) < 0.01
    spec32 = uo.UnetSpec(dim=32, init_dim=32, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, deep_feature=False)
    assert abs(uo.count_flops(spec32, 1, 32) / 1e9 - 46.57) < 0.01


def test_reference_cond_scale_is_an_identity_for_the_3d_unet(ref):
    """The claim behind Imagen.p_sample_loop's handling of cond_scale != 1: the reference's 3-D Unet.forward ignores cond_drop_prob, so
    its classifier-free-guidance double forward (:1540-1552) returns the plain forward bit for bit."""
    case = FORWARD_CASES["deep_dim32_s8"]
    unet = ref.Unet(**unet_kwargs_for_reference(case)).eval()
    fill_module_(unet, seed=case["weight_seed"])
    x, lr, time = build_inputs(case)
    with torch.no_grad():
        plain = unet(x, None, time, lowres_cond_img=lr)
        guided = unet.forward_with_cond_scale(x, None, time, lowres_cond_img=lr, cond_scale=3.0)
    assert torch.equal(plain, guided)
