"""Side measurements of the other BASELINE.json configurations (not bench.py lines): DDPM at larger batches and the Elucidated
(Karras / Heun) sampler at batch B (config 4: 64^3 patches, batch 32 per GPU, 32 steps = 63 U-Net forwards).
    python tools/bench_configs.py [batch ...]
Prints one JSON line per measurement: patches/s, ms per U-Net forward, fraction of the sustained bf16 peak (all conv FLOPs / wall time)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusioniqt_b200 import ElucidatedImagen, NullUnet, SRUnet256
from diffusioniqt_b200.synth import synthetic_field, synthetic_state_dict

dev = torch.device("cuda")
peaks = bench.read_peaks()
S = 64
batches = [int(a) for a in sys.argv[1:]] or [1, 8, 32]


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B in batches:
    # DDPM, T iterations of the driver U-Net at batch B
    T = 20
    imagen = bench.build_model(T, dev, "bf16")
    unet, sched = imagen.unets[1], imagen.noise_schedulers[1]
    imagen.return_host_lists = False
    lr = synthetic_field((B, 1, S, S, S), 5).to(dev)
    ms = timed(lambda: imagen.p_sample_loop(unet, (B, 1, S, S, S), noise_scheduler=sched, lowres_cond_img=lr, pred_objective="x_start",
                                            dynamic_threshold=False, use_tqdm=False), 2)
    per_fwd = ms / T
    tf = bench.FLOPS_PER_FWD_64 * B / (per_fwd * 1e-3) / 1e12
    print(json.dumps(dict(sampler="ddpm", batch=B, timesteps=T, ms_per_forward=per_fwd, ms_per_forward_per_patch=per_fwd / B,
                          patches_per_s_at_T1000=B / (per_fwd * 1000 * 1e-3), tflops=tf, frac_sustained_peak=tf / peaks["sustained"])), flush=True)
    del imagen, unet
    torch.cuda.empty_cache()
    # Elucidated sampler (BASELINE config 4), 32 steps = 63 forwards
    u = SRUnet256(**bench.DRIVER_UNET, img_size=S)
    u.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in u.state_dict().items()}, seed=0))
    edm = ElucidatedImagen(unets=(NullUnet(), u), image_sizes=(S, S), channels=1, cond_drop_prob=0.0, auto_normalize_img=False,
                           dynamic_thresholding=False, num_sample_steps=32).to(dev)
    edm.unets[1].set_compute_dtype("bf16")
    t0 = None
    ms = timed(lambda: edm.sample(batch_size=B, start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False), 1)
    fwd = 63
    tf = bench.FLOPS_PER_FWD_64 * B * fwd / (ms * 1e-3) / 1e12
    print(json.dumps(dict(sampler="elucidated", batch=B, forwards=fwd, ms_per_sample=ms, patches_per_s=B / (ms * 1e-3), tflops=tf,
                          frac_sustained_peak=tf / peaks["sustained"])), flush=True)
    del edm, u
    torch.cuda.empty_cache()
