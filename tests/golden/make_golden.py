"""Generate the committed golden fixtures by running the UNMODIFIED reference here.

    python tests/golden/make_golden.py          # needs /root/reference (read-only)

The reference ships no tests or known-answer vectors, so the pins for
this repo's oracle are outputs of the reference itself: this script imports
/root/reference/imagen_pytorch3D.py through `ref_shim`, builds `Unet` / `Imagen` exactly the way
`train.py:80-133` does (scaled-down sizes), overwrites the parameters with the deterministic
values of `diffusioniqt_b200.synth` (keyed by state_dict name, so no weights need storing), feeds
deterministic inputs and an injected noise sequence, and stores only the outputs.

Fixtures are tiny `.npz` files under tests/golden/.  `cases.py` holds the case table shared with
the tests, which rebuild weights/inputs from the same seeds.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from cases import FORWARD_CASES, SAMPLE_CASES, build_inputs, unet_kwargs_for_reference, make_configs, tap_digest  # noqa: E402
from diffusioniqt_b200.synth import fill_module_, synthetic_noise  # noqa: E402


class _InjectedNoise:
    """Replace torch.randn / torch.randn_like with a recorded sequence while sampling."""

    def __init__(self, seq):
        self.seq = list(seq)
        self.i = 0

    def _next(self, shape):
        v = self.seq[self.i]
        self.i += 1
        assert tuple(v.shape) == tuple(shape), (v.shape, shape)
        return v.clone()

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *size, **kw: self._next(size[0] if len(size) == 1 and not isinstance(size[0], int) else size)
        torch.randn_like = lambda t, **kw: self._next(t.shape)
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def build_reference_unet(ref, case):
    """Reference `Unet` with the case's constructor kwargs and synthetic parameters."""
    unet = ref.Unet(**unet_kwargs_for_reference(case)).eval()
    fill_module_(unet, seed=case["weight_seed"])
    return unet


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_shim.load_reference()

    import json
    only = set(sys.argv[1:])          # optional: regenerate just the named cases (the state_dict contract is always rewritten)
    contract = {}
    for name, case in FORWARD_CASES.items():
        unet = build_reference_unet(ref, case)
        contract[name] = [[k, list(v.shape)] for k, v in unet.state_dict().items()]
        if only and name not in only:
            continue
        x, lr, time = build_inputs(case)
        acts = {}
        hooks = []
        for mod_name in case.get("taps", ()):  # record a few intermediate activations
            mod = unet.get_submodule(mod_name)
            hooks.append(mod.register_forward_hook(lambda m, i, o, n=mod_name: acts.__setitem__(n, o.detach().clone())))
        with torch.no_grad():
            y = unet(x, None, time, lowres_cond_img=lr)
        for h in hooks:
            h.remove()
        out = {"out": y.numpy()}
        for k, v in acts.items():            # a strided digest keeps the fixture small
            out["tap:" + k] = tap_digest(v).numpy()
        np.savez_compressed(os.path.join(HERE, f"fwd_{name}.npz"), **out)
        print(f"fwd_{name}: out {tuple(y.shape)} std {y.std():.4f}")

    with open(os.path.join(HERE, "state_dict_contract.json"), "w") as f:
        json.dump(contract, f)

    for name, case in SAMPLE_CASES.items():
        if only and name not in only:
            continue
        unet = build_reference_unet(ref, case)
        configs = make_configs(case)
        S, B, T = case["size"], case["batch"], case["timesteps"]
        imagen = ref.Imagen(
            unets=(ref.NullUnet(), unet), configs=configs, image_sizes=(S, S), channels=1,
            min_bound=case["min_bound"], timesteps=T, pred_objectives=case.get("pred_objective", "x_start"),
            dynamic_thresholding=case.get("dynamic_threshold", False), p2_loss_weight_gamma=0.0,
            auto_normalize_img=False, cond_drop_prob=0.0, boundary=case.get("boundary", False))
        _, lr, _ = build_inputs(case)
        n_steps = len(range(0, T, case["skip_steps"])) + 1 if case.get("skip_steps") else T
        noise = synthetic_noise((B, 1, S, S, S), n_steps + 1, case["noise_seed"])
        with _InjectedNoise(noise):
            img, traj_x, traj_x0 = imagen.sample(batch_size=B, start_image_or_video=lr, start_at_unet_number=2,
                                                 skip_steps=case.get("skip_steps"), use_tqdm=False)
        keep = sorted(set([0, len(traj_x0) // 2, len(traj_x0) - 1]))
        out = {"img": img.numpy(), "keep": np.array(keep)}
        for k in keep:
            out[f"x_t:{k}"] = traj_x[k]
            out[f"x0:{k}"] = traj_x0[k]
        np.savez_compressed(os.path.join(HERE, f"sample_{name}.npz"), **out)
        print(f"sample_{name}: img std {img.std():.4f} min {img.min():.4f} steps {len(traj_x0) - 1}")


if __name__ == "__main__":
    main()
