"""Golden fixtures for the Elucidated (Karras / Heun) sampler, produced by the reference's OWN loop code.

    python tests/golden/make_golden_elucidated.py          # needs /root/reference (read-only)

`ElucidatedImagen.__init__` cannot be run with the 3-D `Unet` (it passes `cond_on_text` /
`text_embed_dim` to `Unet.cast_model_parameters`, which takes neither), so the instance is assembled with
`__new__` + the handful of attributes `one_unet_sample` reads, and the UNMODIFIED reference methods
`one_unet_sample`, `preconditioned_network_forward`, `sample_schedule`, `c_in/c_out/c_skip/c_noise` and
`threshold_x_start` (elucidated_imagen.py:298-532) are then executed around the reference `Unet` through the
adapter of Appendix C:  unet(c_in * x, <unused>, c_noise, lowres_cond_img=lr), sigma padded to 5-D.
Only the outputs are stored; weights / inputs / noise are regenerated from seeds by the tests.
"""
from __future__ import annotations

import os
import sys
import types
from functools import partial

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from cases import ELUCIDATED_CASES, build_inputs, elucidated_hparams, elucidated_noise_count  # noqa: E402
from diffusioniqt_b200.synth import synthetic_noise  # noqa: E402
from make_golden import _InjectedNoise, build_reference_unet  # noqa: E402


def load_reference_elucidated():
    ref_shim.load_reference()
    if "elucidated_imagen" in sys.modules:
        return sys.modules["elucidated_imagen"]
    if "t5" not in sys.modules:   # t5.py fetches a HuggingFace config at import; never used on this path
        t5 = types.ModuleType("t5")
        t5.DEFAULT_T5_NAME = "none"
        t5.get_encoded_dim = lambda name: 0
        t5.t5_encode_text = lambda *a, **k: None
        sys.modules["t5"] = t5
    sys.path.insert(0, ref_shim.REFERENCE_DIR)
    try:
        import elucidated_imagen  # type: ignore
    finally:
        sys.path.remove(ref_shim.REFERENCE_DIR)
    return elucidated_imagen


class UnetAdapter:
    """What `one_unet_sample` calls: forward_with_cond_scale(x, c_noise, **kwargs) (elucidated_imagen.py:347-351)."""

    self_cond = False

    def __init__(self, unet, lowres):
        self.unet, self.lowres = unet, lowres

    def forward_with_cond_scale(self, x, time, *, cond_scale=1., self_cond=None, **kwargs):
        return self.unet(x, None, time, lowres_cond_img=self.lowres)


def build_reference_sampler(mod, hp, percentile=0.95):
    from einops import rearrange
    E = mod.ElucidatedImagen
    obj = E.__new__(E)
    nn.Module.__init__(obj)
    obj.hparams = [mod.Hparams(**hp)]
    obj.register_buffer("_temp", torch.tensor([0.]), persistent=False)
    obj.right_pad_dims_to_datatype = partial(rearrange, pattern="b -> b 1 1 1 1")      # the is_video pattern of :165
    obj.dynamic_thresholding_percentile = percentile
    obj.normalize_img = obj.unnormalize_img = lambda t: t                              # auto_normalize_img = False
    return obj


def run_reference(case):
    mod = load_reference_elucidated()
    ref = ref_shim.load_reference()
    unet = build_reference_unet(ref, case)
    _, lr, _ = build_inputs(case)
    B, S = case["batch"], case["size"]
    hp = elucidated_hparams(case)
    noise = synthetic_noise((B, 1, S, S, S), elucidated_noise_count(case), case["noise_seed"])
    sampler = build_reference_sampler(mod, hp)
    with torch.no_grad(), _InjectedNoise(noise) as inj:
        img = sampler.one_unet_sample(UnetAdapter(unet, lr), (B, 1, S, S, S), unet_number=1, clamp=True,
                                      dynamic_threshold=case["dynamic_threshold"], cond_scale=1., use_tqdm=False,
                                      skip_steps=case.get("skip_steps"))
        assert inj.i == len(noise), (inj.i, len(noise))
    return img


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, case in ELUCIDATED_CASES.items():
        img = run_reference(case)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), img=img.numpy())
        print(f"{name}: img {tuple(img.shape)} std {img.std():.4f} min {img.min():.4f} max {img.max():.4f}")


if __name__ == "__main__":
    main()
