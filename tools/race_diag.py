"""Race diagnostics: the volume acceptance case sampled patch by patch, compared with a reference pass of the SAME GPU code run with
CUDA graphs off and a host sync per sample.  Variant switches (env): NOISE=cpu|gpu|pinned, DIQT_DISABLE_GROUPED, DIQT_DISABLE_PDL,
DIQT_DEBUG_SYNC_REPLAY=before|after|both, HOSTLISTS=0|1, DTYPE."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from cases import MIN_BOUND
from diffusioniqt_b200 import Imagen, NullUnet, Unet, volume as V
from diffusioniqt_b200.synth import synthetic_field, synthetic_noise, synthetic_state_dict
from test_gpu_volume import KW

dtype = os.environ.get("DTYPE", "fp32")
noise_mode = os.environ.get("NOISE", "cpu")
P, stride, T, N = 16, 8, 6, 32
unet = Unet(**KW, img_size=P)
unet.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=61))
imagen = Imagen(unets=(NullUnet(), unet), configs={"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}, image_sizes=(P, P), channels=1,
                min_bound=MIN_BOUND, timesteps=T, pred_objectives="x_start", dynamic_thresholding=False, cond_drop_prob=0.0).cuda()
imagen.unets[1].set_compute_dtype(dtype)
imagen.return_host_lists = os.environ.get("HOSTLISTS", "1") == "1"
lowres = synthetic_field((N, N, N), 62)[...]
grid = V.patch_grid(lowres.shape, P, stride)
noise = {g: synthetic_noise((1, 1, P, P, P), T + 1, 64 + n) for n, g in enumerate(grid)}
if noise_mode == "gpu":
    noise = {g: [t.cuda() for t in v] for g, v in noise.items()}
elif noise_mode == "pinned":
    noise = {g: [t.pin_memory() for t in v] for g, v in noise.items()}


def run_all(sync):
    outs = []
    for g in grid:
        lr = lowres[g[0]:g[0] + P, g[1]:g[1] + P, g[2]:g[2] + P][None, None].cuda()
        imagen.noise_override = noise[g]
        outs.append(imagen.sample(batch_size=1, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0])
        if sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return outs


graph = imagen.use_cuda_graph
imagen.use_cuda_graph = False
ref = run_all(True)
imagen.use_cuda_graph = graph
bad_total = []
for rep in range(3):
    outs = run_all(False)
    bad = [i for i, (a, b) in enumerate(zip(outs, ref)) if not torch.equal(a, b)]
    bad_total.append(bad)
tag = " ".join(f"{k}={os.environ[k]}" for k in ("DTYPE", "NOISE", "HOSTLISTS", "DIQT_DISABLE_GROUPED", "DIQT_DISABLE_PDL", "DIQT_DEBUG_SYNC_REPLAY", "DIQT_DISABLE_CUDA_GRAPH") if k in os.environ)
print(f"[{tag or 'baseline'}] bad patches per repeat: {bad_total}")
