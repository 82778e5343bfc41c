"""Host-side mirror of the reference API: state_dict contract, schedule table, error behaviour (CPU only)."""
import json
import os

import pytest
import torch

from cases import FORWARD_CASES, MIN_BOUND, unet_kwargs_for_reference
from helpers import GOLDEN_DIR
from oracle import ddpm_oracle as do


@pytest.mark.parametrize("name", list(FORWARD_CASES))
def test_state_dict_contract(name):
    from diffusioniqt_b200 import Unet
    contract = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_contract.json")))[name]
    kw = unet_kwargs_for_reference(FORWARD_CASES[name])
    got = [[k, list(v.shape)] for k, v in Unet(**kw).state_dict().items()]
    assert got == contract


def test_aliases_and_presets():
    import diffusioniqt_b200 as pkg
    assert pkg.Unet3D is pkg.Unet
    u = pkg.SRUnet256(dim=32, init_dim=32, dim_mults=(1, 2), num_resnet_blocks=(1, 1), channels=1, lowres_cond=True, init_cross_embed=False,
                      attend_at_middle=False, attend_at_enc=(False, False), memory_efficient=False, deep_feature=False)
    assert isinstance(u, pkg.Unet)
    v = u.cast_model_parameters(lowres_cond=True, channels=1, channels_out=1)
    assert v is u
    w = u.cast_model_parameters(lowres_cond=False, channels=1, channels_out=1)
    assert w is not u and w.init_conv.weight.shape[1] == 1 and type(w) is pkg.SRUnet256


def test_unsupported_options_raise():
    from diffusioniqt_b200 import Unet
    base = dict(dim=32, init_dim=32, dim_mults=(1, 2), channels=1, lowres_cond=True, init_cross_embed=False, attend_at_middle=False,
                attend_at_enc=(False, False), deep_feature=False)
    for bad in (dict(init_cross_embed=True, boundary=True), dict(memory_efficient=True), dict(pixel_shuffle_upsample=False, boundary=True), dict(attend_at_enc=(True, False), attn_dim_head=48),
                dict(self_cond=True), dict(cross_embed_downsample=True)):
        with pytest.raises(NotImplementedError):
            Unet(**{**base, **bad})


def test_cpu_tensors_fail_loudly():
    from diffusioniqt_b200 import Unet
    u = Unet(dim=32, init_dim=32, dim_mults=(1, 2), channels=1, lowres_cond=True, init_cross_embed=False, attend_at_middle=False,
             attend_at_enc=(False, False), deep_feature=False)
    x = torch.zeros(1, 1, 8, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        u(x, None, torch.zeros(1), lowres_cond_img=x)


@pytest.mark.parametrize("T,skip", [(12, None), (20, 4), (1000, None)])
def test_schedule_table_matches_oracle(T, skip):
    from diffusioniqt_b200 import Imagen, NullUnet, Unet
    u = Unet(dim=32, init_dim=32, dim_mults=(1, 2), channels=1, lowres_cond=True, init_cross_embed=False, attend_at_middle=False,
             attend_at_enc=(False, False), deep_feature=False)
    im = Imagen(unets=(NullUnet(), u), configs={"Data": {"norm": "z-score"}}, image_sizes=(8, 8), channels=1, timesteps=T,
                pred_objectives="x_start", dynamic_thresholding=False, min_bound=MIN_BOUND, cond_drop_prob=0.0)
    table, log_snr = im._build_schedule(im.noise_schedulers[1], skip, "x_start", torch.device("cpu"))
    pairs = do.sampling_timesteps(T, skip)
    assert table.shape == (len(pairs), 8)
    x_t, x0, eps = torch.tensor([0.3]), torch.tensor([-0.2]), torch.tensor([1.7])
    for i, (t, tn) in enumerate(pairs):
        tt, ttn = torch.tensor([t]), torch.tensor([tn])
        mean, _, log_var = do.q_posterior(x0, x_t, tt, ttn)
        want = mean + (0.0 if tn == 0 else 1.0) * (0.5 * log_var).exp() * eps
        al, _, c, an, ns, lo, hi, obj = table[i].tolist()
        got = an * (0.3 * (1 - c) / al + c * -0.2) + ns * 1.7
        assert abs(got - want.item()) < 1e-5 * max(1.0, abs(want.item()))
        assert abs(log_snr[i].item() - do.alpha_cosine_log_snr(tt).item()) < 1e-5
        assert lo == pytest.approx(MIN_BOUND) and hi == float("inf") and obj == 0
    assert table[-1, 4].item() == 0.0      # no noise on the last step (:2053-2054)


def test_noop_moves_keep_engines_real_changes_drop_them():
    """Imagen.sample calls `.to(device)` on every call (imagen_pytorch3D.py:1941-1962): a no-op move must not drop the compiled
    engines (and the CUDA graphs cached on them); a real change of the parameters must."""
    from diffusioniqt_b200 import Unet

    class _Eng:
        closed = False

        def close(self):
            self.closed = True

    u = Unet(dim=32, init_dim=32, dim_mults=(1, 2), channels=1, lowres_cond=True, init_cross_embed=False, attend_at_middle=False,
             attend_at_enc=(False, False), deep_feature=False)
    e = _Eng()
    u._engines["k"] = e
    torch.nn.ModuleList([u]).to("cpu")
    u.float()
    assert u._engines == {"k": e} and not e.closed
    u.double()
    assert u._engines == {} and e.closed
    e2 = _Eng()
    u._engines["k"] = e2
    u.load_state_dict(u.state_dict())
    assert e2.closed and u._engines == {}
