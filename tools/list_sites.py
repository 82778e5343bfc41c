"""Print the conv sites of the config-2 engine and which of them carry the fused GroupNorm (diagnostics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusioniqt_b200 import Unet, lib as L
from diffusioniqt_b200.synth import synthetic_state_dict
S = 64
kw = dict(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, init_cross_embed=False, attend_at_middle=False,
          attend_at_enc=(False, False, False), use_se_attn=True, memory_efficient=False, deep_feature=False, boundary=False, batch_sample=False, img_size=S)
unet = Unet(**kw)
unet.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=5))
unet = unet.cuda().set_compute_dtype("bf16")
x = torch.randn(1, 1, S, S, S, device="cuda")
unet(x, None, torch.tensor([1.0], device="cuda"), lowres_cond_img=x)
eng = next(iter(unet._engines.values()))
names = {L.IMPL_ZM: "zm", L.IMPL_TC: "tc", L.IMPL_SIMT: "simt"}
for k, v in eng.conv_impls.items():
    print(f"{k:45s} {names.get(v, v):5s} {'fused-gn' if k in eng.fused_gn else ''}")
