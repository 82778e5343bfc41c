"""BASELINE config 3: a whole synthetic 256^3 low-field volume as overlapping 64^3 patches (stride 32, crop margin 16 -> 7^3 = 343
patches), sharded over the ranks, one all-gather, device-side stitch.  T denoising steps per patch (default 20, the eval_config.yaml
setting; the metric config of bench.py uses 1000).  Run plain (1 GPU) or under torchrun (N ranks).  Prints one JSON line on rank 0."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusioniqt_b200 import volume as V
from diffusioniqt_b200.synth import synthetic_field

T = int(os.environ.get("T", "20"))
BATCH = int(os.environ.get("BATCH", "7"))
N = int(os.environ.get("SIDE", "256"))
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
imagen = bench.build_model(T, dev, "bf16")
imagen.return_host_lists = False
low = synthetic_field((N, N, N), 11)
low[: N // 8] = low.min()                                   # air: some patches fall under the 5 % rule
low = low.to(dev)
raw = low - low.min()


def sample_fn(lr):
    return imagen.sample(batch_size=lr.shape[0], start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0]


def run():
    return V.infer_volume(sample_fn, low, patch=64, overlap=32, raw_lowres=raw, batch_size=BATCH, fill_value=bench.MIN_BOUND, rank=rank, world=world)


torch.manual_seed(rank)
warm = V.infer_volume(sample_fn, low[:96, :96, :96].contiguous(), patch=64, overlap=32, raw_lowres=raw[:96, :96, :96].contiguous(), batch_size=BATCH,
                      fill_value=bench.MIN_BOUND, rank=rank, world=world)     # engine build + graph capture for both batch sizes in use
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = run()
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
dt = time.perf_counter() - t0
if dist is not None:
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t)
vol = res.volume
ok = bool(torch.isfinite(vol).all()) and float(vol.min()) >= min(bench.MIN_BOUND, float(low.min())) - 1e-5
if rank == 0:
    print(json.dumps(dict(config="BASELINE config 3: 256^3 volume, 64^3 patches, stride 32", side=N, timesteps=T, batch_per_call=BATCH, n_gpus=world,
                          patches=res.n_patches, skipped=res.n_skipped, patches_per_rank=res.patches_per_rank, seconds=dt,
                          patches_per_s=res.n_patches / dt, finite_and_bounded=ok)), flush=True)
if dist is not None:
    dist.destroy_process_group()
