"""Training step at the BASELINE config-2 shape: one 64^3 patch, driver U-Net (dim 64), bf16: this library's forward-with-tape +
reverse pass against PyTorch autograd over the reference's op list (oracle.unet_forward = the ATen / cuDNN calls of imagen_pytorch3D.py:
1554-1684) on the same GPU.  Host-timed around a device synchronize (both sides launch hundreds of kernels from Python)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import torch.nn.functional as F
from diffusioniqt_b200 import Unet, lib as L
from diffusioniqt_b200.synth import synthetic_field, synthetic_state_dict
from diffusioniqt_b200.train import UnetBackprop
from oracle.unet_oracle import UnetSpec, unet_forward

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = dict(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, init_cross_embed=False, attend_at_middle=False,
          attend_at_enc=(False, False, False), use_se_attn=True, memory_efficient=False, deep_feature=False, boundary=False, batch_sample=False, img_size=S)
unet = Unet(**kw)
sd = synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=5)
unet.load_state_dict(sd)
unet = unet.cuda().set_compute_dtype("bf16")
spec = UnetSpec(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, deep_feature=False)
x, lr = synthetic_field((B, 1, S, S, S), 3).cuda(), synthetic_field((B, 1, S, S, S), 4).cuda()
t = torch.full((B,), 1.3, device="cuda")
target = torch.randn_like(x)


def ours():
    bp = UnetBackprop(unet)
    pred = bp.forward(x, t, lowres_cond_img=lr)
    bp.backward(2 * (pred - target) / pred.numel())
    for p in unet.parameters():
        p.grad = None


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


n0 = L.launch_count()
ours()
launches = L.launch_count() - n0
if os.environ.get("BENCH_TRAIN_ONLY"):
    sys.exit(0)
res = dict(op="training step (forward with tape + reverse pass)", patch=S, batch=B, dtype="bf16", ours_ms=timed(ours, 5), ours_kernel_launches=launches)

# the whole step (forward with tape, loss gradient, reverse pass, Adam) captured once as a CUDA graph and replayed: the kernels are ordered
# by the stream alone, so one graph launch per optimizer step
from diffusioniqt_b200.train import AdamState
opt = AdamState(unet.parameters(), lr=1e-6)


def full_step():
    bp = UnetBackprop(unet)
    pred = bp.forward(x, t, lowres_cond_img=lr)
    bp.backward(2 * (pred - target) / pred.numel())
    opt.step()
    opt.zero_grad()


try:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            full_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    res["ours_step_with_adam_ms"] = timed(full_step, 5)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        full_step()
    res["ours_graph_replay_ms"] = timed(g.replay, 10)
except Exception as e:  # noqa: BLE001
    res["ours_graph_error"] = repr(e)[:300]
    torch.cuda.synchronize()
sdg = {k: v.cuda().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
ref_opt = torch.optim.Adam([v for v in sdg.values() if v.requires_grad], lr=1e-6, fused=True)
for name, ctx, tf32 in (("torch_fp32_tf32", None, True), ("torch_bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16), True)):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32

    def ref():
        if ctx is not None:
            with ctx:
                out = unet_forward(sdg, spec, x, t, lowres_cond_img=lr)
        else:
            out = unet_forward(sdg, spec, x, t, lowres_cond_img=lr)
        F.mse_loss(out.float(), target).backward()
        ref_opt.step()
        ref_opt.zero_grad(set_to_none=True)
    try:
        res[name + "_ms"] = timed(ref, 5)
    except Exception as e:  # noqa: BLE001
        res[name + "_ms"] = None
        res[name + "_error"] = str(e)[:200]
print(json.dumps(res))
