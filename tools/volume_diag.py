"""Per-patch diagnostics of the volume acceptance case (tests/test_gpu_volume.py): GPU sampler vs CPU oracle per patch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
from cases import MIN_BOUND
from diffusioniqt_b200 import Imagen, NullUnet, Unet, volume as V
from diffusioniqt_b200.synth import synthetic_field, synthetic_noise, synthetic_state_dict
from helpers import spec_from_kwargs
from oracle import stitch_oracle as so
from oracle.ddpm_oracle import ddpm_sample
from oracle.unet_oracle import unet_forward
from test_gpu_volume import KW

dtype = sys.argv[1] if len(sys.argv) > 1 else "fp32"
sync = len(sys.argv) > 2 and sys.argv[2] == "sync"
P, stride, T, N = 16, 8, 6, 32
unet = Unet(**KW, img_size=P)
sd = synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=61)
unet.load_state_dict(sd)
imagen = Imagen(unets=(NullUnet(), unet), configs={"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}, image_sizes=(P, P), channels=1,
                min_bound=MIN_BOUND, timesteps=T, pred_objectives="x_start", dynamic_thresholding=False, cond_drop_prob=0.0).cuda()
imagen.unets[1].set_compute_dtype(dtype)
lowres = synthetic_field((N, N, N), 62)[...]
lowres[:6, :10] = lowres.min()
grid = V.patch_grid(lowres.shape, P, stride)
raw = (lowres - lowres.min()).numpy()
kept = [g for g in grid if not so.is_skipped(raw, list(g), P)]
noise = {g: synthetic_noise((1, 1, P, P, P), T + 1, 64 + n) for n, g in enumerate(grid)}
spec = spec_from_kwargs(KW)
outs = []
for g in kept:      # all GPU work first, no host sync in between (like infer_volume)
    lr = lowres[g[0]:g[0] + P, g[1]:g[1] + P, g[2]:g[2] + P][None, None].cuda()
    imagen.noise_override = noise[g]
    outs.append(imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0])
    if sync:
        torch.cuda.synchronize()
errs = []
for g, o in zip(kept, outs):
    lr = lowres[g[0]:g[0] + P, g[1]:g[1] + P, g[2]:g[2] + P][None, None]
    with torch.no_grad():
        want, _, _ = ddpm_sample(lambda x, ls: unet_forward(sd, spec, x, ls, lowres_cond_img=lr), (1, 1, P, P, P), noise[g], timesteps=T, min_bound=MIN_BOUND)
    e = ((o.cpu() - want).norm() / want.norm()).item()
    errs.append(e)
print(dtype, "sync" if sync else "nosync", "PDL off" if os.environ.get("DIQT_DISABLE_PDL") == "1" else "PDL on", "graph off" if os.environ.get("DIQT_DISABLE_CUDA_GRAPH") == "1" else "graph on")
print("per-patch rel-L2:", " ".join(f"{e:.2e}" for e in errs))
print("max", max(errs), "median", float(np.median(errs)))
