#!/usr/bin/env python
"""Benchmark of the DiffusionIQT sampling hot path (BASELINE.json metric: 3-D patches/sec through the
full denoise loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full pass of the sampler (T denoising iterations of the driver-config U-Net) over one
batch of synthetic low-field patches.  Workload at N=1 is BASELINE config 2: a single 64^3 patch, the U-Net of
train.py:83-116 + config/config.yaml, T = 1000, bf16.  For N > 1 (torchrun, one rank per GPU) every rank
denoises its own patch(es) (weak scaling) and the denoised patches are all-gathered for stitching.

Prints ONE JSON line on rank 0 (see the task contract): value = device-resident throughput, e2e = the same
through Imagen.sample() with host buffers, roofline = the dominant kernel (3x3x3 conv 64->64 at 64^3) timed
live with CUDA events, cpu_baseline = the CPU oracle port timed on this box's host cores.
`--impl reference` times the reference's CPU algorithm (oracle port; the Python reference itself cannot
travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DRIVER_UNET = dict(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True,
                   init_cross_embed=False, attend_at_middle=False, attend_at_enc=(False, False, False), use_se_attn=True,
                   memory_efficient=False, pixel_shuffle_upsample=True, deep_feature=False, boundary=False, batch_sample=False)
MIN_BOUND = (0.0 - 271.64814106698583) / 377.117173547721
FLOPS_PER_FWD_64 = 1488.442865152e9       # BASELINE.md section 3 (B=1, 64^3, driver config)


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p.get("bf16_tflops", 1590.0), sustained=p.get("bf16_tflops_sustained", 1400.0), hbm=p.get("hbm_gbs", 6650.0), source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


def kernel_source_sha():
    """Hash of the sources the dominant kernel is built from: ties a committed ncu capture to the code it profiled."""
    import hashlib
    h = hashlib.sha256()
    for name in ("conv_zm.cu", "tc_common.cuh", "common.cuh"):
        with open(os.path.join(ROOT, "diffusioniqt_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def read_traffic(kernel_label):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json, written by
    tools/update_traffic.py from an `ncu --set full` CSV).  None when there is no capture or when the kernel's sources changed since
    it was taken (a stale figure is worse than none)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        for k, v in t.items():
            if k.split()[0] in kernel_label and "64^3" in kernel_label:
                if v.get("source_sha") != kernel_source_sha():
                    return None, "capture %s is older than the kernel sources (sha %s != %s): not reported" % (v.get("source", "?"), v.get("source_sha"), kernel_source_sha())
                return v["bytes_per_launch"], "dram__bytes_read.sum + dram__bytes_write.sum per launch, %s" % v.get("source", "?")
    except (OSError, ValueError, KeyError):
        pass
    return None, "no ncu capture committed"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                    power_w_max=max(power) if power else None)


def build_model(timesteps, device, dtype):
    from diffusioniqt_b200 import Imagen, NullUnet, SRUnet256
    from diffusioniqt_b200.synth import synthetic_state_dict
    unet = SRUnet256(**DRIVER_UNET, img_size=64)
    unet.load_state_dict(synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=0))
    configs = {"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}
    imagen = Imagen(unets=(NullUnet(), unet), configs=configs, image_sizes=(64, 64), channels=1, min_bound=MIN_BOUND, timesteps=timesteps,
                    pred_objectives="x_start", dynamic_thresholding=False, p2_loss_weight_gamma=0.0, auto_normalize_img=False,
                    cond_drop_prob=0.0).to(device)
    imagen.unets[1].set_compute_dtype(dtype)
    return imagen


def time_dominant_kernel(size, batch, reps=20, kinds=("fused", "plain")):
    """The 3x3x3 conv 64->64 at full resolution (82 % of all FLOPs), alone, exactly as the engine launches
    it in the step: conv_zm_kernel with GroupNorm + FiLM + Mish of its input fused into the load path and the output statistics in its
    epilogue (`fused`), and the same conv without the fused normalisation (`plain`: the kernel the r1 roofline line described).
    `reps` launches over 4 rotating buffer sets (4 x 2 x 32 MiB at 64^3: consecutive launches never hit the same L2 lines) captured as
    ONE CUDA graph -- back to back on the stream with programmatic dependent launch, as inside a sampler iteration -- and CUDA events on
    that stream around 3 replays."""
    import ctypes as C
    from diffusioniqt_b200 import lib as L
    lib = L.load()
    dev = torch.device("cuda")
    n, c = batch, 64
    vox = size ** 3
    nb = 4
    xs = [torch.randn(n * vox, c, device=dev).bfloat16() for _ in range(nb)]
    ys = [torch.empty(n * vox, c, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    w = (torch.randn(c, c, 3, 3, 3, device=dev) * 0.02).bfloat16().float().contiguous()
    b = torch.zeros(c, device=dev)
    gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
    film = (torch.randn(n, 2 * c, device=dev) * 0.2).contiguous()
    desc = L.ConvDesc(mode=L.CONV_K3, dtype=L.BF16, impl=L.IMPL_AUTO, n=n, d0=size, d1=size, d2=size, c_in=c, ld_in=c, c_out=c, ld_out=c, flags=0)
    impl = C.c_int(0)
    L.check(lib.diqt_conv_resolved_impl(C.byref(desc), C.byref(impl)))
    nbytes = C.c_size_t(0)
    L.check(lib.diqt_conv_packed_bytes(C.byref(desc), C.byref(nbytes)))
    packed = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    pb = torch.empty(c, dtype=torch.float32, device=dev)
    st = L.current_stream()
    L.check(lib.diqt_conv_pack(C.byref(desc), w.data_ptr(), b.data_ptr(), packed.data_ptr(), pb.data_ptr(), st))
    fusable = n <= 2 and bool(lib.diqt_conv_gn_fusable(C.byref(desc)))
    # statistics of the inputs (what the producer of x would have left behind) and sinks for the conv's own output statistics
    nblk = max(1, min(vox // 128, 148 // n))
    ng = C.c_int(0)
    L.check(lib.diqt_stats_groups(nblk, 1, C.byref(ng)))
    part = torch.zeros(n * nblk * c * 2, device=dev)
    grp = torch.zeros(16 * n * c * 2, device=dev)
    tick = torch.zeros(16 * n, dtype=torch.int32, device=dev)
    L.check(lib.diqt_channel_stats_g(xs[0].data_ptr(), L.BF16, n, vox, c, c, nblk, part.data_ptr(), grp.data_ptr(), tick.data_ptr(), st))
    opart, ogrp, otick = torch.zeros(n * 320 * c * 2, device=dev), torch.zeros(16 * n * c * 2, device=dev), torch.zeros(16 * n, dtype=torch.int32, device=dev)
    plans = {"fused": [], "plain": []}
    for x, y in zip(xs, ys):
        for kind in kinds:
            if kind == "fused" and not fusable:
                continue
            p = C.c_void_p(0)
            L.check(lib.diqt_conv_plan_create(C.byref(desc), x.data_ptr(), y.data_ptr(), packed.data_ptr(), pb.data_ptr(), C.byref(p)))
            nbk, ngo = C.c_int(0), C.c_int(0)
            L.check(lib.diqt_conv_plan_set_stats_g(p.value, opart.data_ptr(), ogrp.data_ptr(), otick.data_ptr(), C.byref(nbk), C.byref(ngo)))
            if kind == "fused":
                L.check(lib.diqt_conv_plan_set_gn(p.value, grp.data_ptr(), ng.value, vox, 8, 1e-5, gamma.data_ptr(), beta.data_ptr()))
                L.check(lib.diqt_conv_plan_set_film(p.value, film.data_ptr(), 2 * c, 0, 1))
            plans[kind].append(p.value)

    def graph_ms(pl):
        for p in pl:
            L.check(lib.diqt_conv_run(p, L.current_stream()))
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                L.check(lib.diqt_conv_run(pl[i % len(pl)], L.current_stream()))
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (3 * reps)

    flops = 2.0 * c * c * 27 * n * vox
    out = dict(flops=flops, kernel={L.IMPL_SIMT: "conv_simt_kernel", L.IMPL_TC: "conv_tc_kernel", L.IMPL_ZM: "conv_zm_kernel"}[impl.value])
    for kind, pl in plans.items():
        if pl:
            ms = graph_ms(pl)
            out[kind] = dict(ms=ms, tflops=flops / (ms * 1e-3) / 1e12)
    for pl in plans.values():
        for p in pl:
            lib.diqt_conv_plan_destroy(p)
    main = out.get("fused") or out["plain"]
    out.update(ms=main["ms"], tflops=main["tflops"], as_launched="fused GroupNorm+FiLM+Mish input path" if "fused" in out else "plain")
    return out


def time_elementwise_kernel(size, batch, reps=20):
    """The bandwidth-bound companion of the dominant kernel, as the step launches it: the SE / residual pass of a ResnetBlock
    (out = h * sigmoid(W2 relu(W1 mean(h))) + x, plus the channel statistics of `out` for the next GroupNorm; SE3D :617-632, :612) over a
    64-channel full-resolution tensor: 2 reads + 1 write, alone, rotating buffers (each 32 MiB at 64^3), graph replay, CUDA events."""
    import ctypes as C
    from diffusioniqt_b200 import lib as L
    lib = L.load()
    dev = torch.device("cuda")
    n, c = batch, 64
    vox = size ** 3
    nb = 4
    hs = [torch.randn(n * vox, c, device=dev).bfloat16() for _ in range(nb)]
    xs = [torch.randn(n * vox, c, device=dev).bfloat16() for _ in range(nb)]
    outs = [torch.empty(n * vox, c, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    hidden = c // 16
    w1, w2 = torch.randn(hidden, c, device=dev) * 0.1, torch.randn(c, hidden, device=dev) * 0.1
    nblk = max(1, min(vox // 128, 148 // n))
    ng = C.c_int(0)
    L.check(lib.diqt_stats_groups(nblk, 1, C.byref(ng)))
    part, grp, tick = torch.zeros(n * nblk * c * 2, device=dev), torch.zeros(16 * n * c * 2, device=dev), torch.zeros(16 * n, dtype=torch.int32, device=dev)
    L.check(lib.diqt_channel_stats_g(hs[0].data_ptr(), L.BF16, n, vox, c, c, nblk, part.data_ptr(), grp.data_ptr(), tick.data_ptr(), L.current_stream()))
    opart, ogrp, otick = torch.zeros(n * nblk * c * 2, device=dev), torch.zeros(16 * n * c * 2, device=dev), torch.zeros(16 * n, dtype=torch.int32, device=dev)

    def run(i, st):
        i %= nb
        L.check(lib.diqt_scale_residual_g(hs[i].data_ptr(), c, xs[i].data_ptr(), c, outs[i].data_ptr(), c, L.BF16, n, vox, c, grp.data_ptr(), ng.value, hidden,
                                          w1.data_ptr(), w2.data_ptr(), nblk, opart.data_ptr(), ogrp.data_ptr(), otick.data_ptr(), st))

    for i in range(nb):
        run(i, L.current_stream())
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            run(i, L.current_stream())
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (3 * reps)
    nbytes = 3.0 * n * vox * c * 2
    return dict(ms=ms, bytes=nbytes, gbs=nbytes / (ms * 1e-3) / 1e9)


def cpu_baseline(size, timesteps, batch, denoise_steps, threads=None, warm=True):
    """The CPU oracle port (same algorithm as the reference's CPU sampler) on a bounded sample: `denoise_steps`
    iterations of the sampler at the benchmark shape; patches/s is extrapolated to `timesteps` iterations."""
    from diffusioniqt_b200 import SRUnet256
    from diffusioniqt_b200.synth import synthetic_field, synthetic_state_dict
    from oracle.ddpm_oracle import alpha_cosine_log_snr, q_posterior
    from oracle.unet_oracle import UnetSpec, unet_forward
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    shapes = {k: tuple(v.shape) for k, v in SRUnet256(**DRIVER_UNET, img_size=size).state_dict().items()}
    sd = synthetic_state_dict(shapes, seed=0)
    spec = UnetSpec(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, deep_feature=False)
    lr = synthetic_field((batch, 1, size, size, size), 1)
    x = torch.randn(batch, 1, size, size, size)
    times = torch.linspace(1.0, 0.0, timesteps + 1)

    def one(i):
        nonlocal x
        t, tn = times[i].expand(batch), times[i + 1].expand(batch)
        with torch.no_grad():
            x0 = unet_forward(sd, spec, x, alpha_cosine_log_snr(t), lowres_cond_img=lr).clamp(min=MIN_BOUND)
            mean, _, log_var = q_posterior(x0, x, t, tn)
            x = mean + (0.5 * log_var).exp() * torch.randn_like(x)

    if warm:
        one(0)  # warm-up (oneDNN primitive creation)
    per = []
    for i in range(denoise_steps):
        t0 = time.perf_counter()
        one(1 + i)
        per.append(time.perf_counter() - t0)
    dt = statistics.median(per)
    return dict(value=batch / (dt * timesteps), unit="patches/s", cores=threads, kind="port", ms_per_denoise_step=dt * 1e3,
                ms_per_denoise_step_min=min(per) * 1e3, ms_per_denoise_step_max=max(per) * 1e3,
                sample=f"{denoise_steps} of {timesteps} denoising iterations of one {size}^3 patch (batch {batch}), oracle/ CPU port of the reference "
                       f"sampler, torch {torch.__version__} fp32, median iteration extrapolated linearly to {timesteps} iterations")


def torch_gpu_baseline(size, batch, reps=5):
    """The kernel to beat on the same box: the reference's own op list (oracle.unet_forward = the ATen / cuDNN
    calls of imagen_pytorch3D.py:1554-1684) run eagerly on THIS GPU, one forward = one denoising iteration at the benchmark shape:
    fp32 with TF32 off (the parity-grade reference), fp32 with TF32 on (what the reference does on a GPU by default) and bf16 autocast
    over channels_last_3d tensors.  A reported baseline like cpu_baseline: the oracle is timed, never shipped."""
    from diffusioniqt_b200 import SRUnet256
    from diffusioniqt_b200.synth import synthetic_field, synthetic_state_dict
    from oracle.unet_oracle import UnetSpec, unet_forward
    dev = torch.device("cuda")
    shapes = {k: tuple(v.shape) for k, v in SRUnet256(**DRIVER_UNET, img_size=size).state_dict().items()}
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(shapes, seed=0).items()}
    spec = UnetSpec(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, deep_feature=False)
    x = synthetic_field((batch, 1, size, size, size), 2).to(dev)
    lr = synthetic_field((batch, 1, size, size, size), 1).to(dev)
    t = torch.full((batch,), 1.3, device=dev)
    out = {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True

    def run(mode):
        with torch.no_grad():
            if mode == "bf16_autocast_channels_last_3d":
                sd_ = {k: (v.contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v) for k, v in sd.items()}
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return unet_forward(sd_, spec, x.contiguous(memory_format=torch.channels_last_3d), t,
                                        lowres_cond_img=lr.contiguous(memory_format=torch.channels_last_3d)).float()
            return unet_forward(sd, spec, x, t, lowres_cond_img=lr)

    try:
        for mode, tf32 in (("fp32_tf32_off", False), ("fp32_tf32_on", True), ("bf16_autocast_channels_last_3d", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                for _ in range(2):
                    run(mode)
                torch.cuda.synchronize()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
                ev[0].record()
                for i in range(reps):
                    run(mode)
                    ev[i + 1].record()
                torch.cuda.synchronize()
                ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
                out[mode] = dict(ms_per_denoise_iteration=ms[len(ms) // 2], ms_min=ms[0], ms_max=ms[-1])
            except Exception as e:  # a cuDNN configuration that does not run is a fact worth reporting, not a bench failure
                out[mode] = dict(error=f"{type(e).__name__}: {e}"[:200])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["note"] = ("eager PyTorch %s / cuDNN %s forward of the reference op list on this GPU, batch %d, %d^3, median of %d; sampler update excluded"
                   % (torch.__version__, torch.backends.cudnn.version(), batch, size, reps))
    del sd
    torch.cuda.empty_cache()
    return out


def train_step_record(size, batch, reps=5):
    """SURVEY 8 f-4 next to the headline: one optimizer step (forward with the activations kept, loss gradient, reverse pass, Adam) of the
    driver U-Net on one patch, bf16, captured once and replayed as ONE CUDA graph (CUDA events around the replays), against PyTorch autograd
    over the reference's op list (oracle.unet_forward, the ATen / cuDNN calls of imagen_pytorch3D.py:1554-1684) + torch.optim.Adam(fused) on
    the same GPU.  A baseline leg like torch_gpu_baseline: the oracle is timed, never shipped."""
    import torch.nn.functional as F
    from diffusioniqt_b200 import SRUnet256, lib as L
    from diffusioniqt_b200.synth import synthetic_field, synthetic_state_dict
    from diffusioniqt_b200.train import AdamState, UnetBackprop
    from oracle.unet_oracle import UnetSpec, unet_forward
    dev = torch.device("cuda")
    unet = SRUnet256(**DRIVER_UNET, img_size=size)
    sd = synthetic_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=0)
    unet.load_state_dict(sd)
    unet = unet.to(dev).set_compute_dtype("bf16")
    x = synthetic_field((batch, 1, size, size, size), 2).to(dev)
    lr = synthetic_field((batch, 1, size, size, size), 1).to(dev)
    t = torch.full((batch,), 1.3, device=dev)
    target = synthetic_field((batch, 1, size, size, size), 3).to(dev)
    opt = AdamState(unet.parameters(), lr=1e-6)
    out = {}

    def step():
        bp = UnetBackprop(unet)
        pred = bp.forward(x, t, lowres_cond_img=lr)
        bp.backward(2 * (pred - target) / pred.numel())
        opt.step()
        opt.zero_grad()

    def event_ms(fn):
        fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        return ms[len(ms) // 2]

    n0 = L.launch_count()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
        launches = L.launch_count() - n0
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    out["ms_launched_from_python"] = event_ms(step)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        out["ms_graph_replay"] = event_ms(g.replay)
    except Exception as e:  # noqa: BLE001
        out["graph_error"] = f"{type(e).__name__}: {e}"[:200]
        torch.cuda.synchronize()
    out["kernel_launches_of_this_library"] = int(launches)
    sdg = {k: v.to(dev).requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    spec = UnetSpec(dim=64, init_dim=64, dim_mults=(1, 2, 4), num_resnet_blocks=(2, 2, 2), channels=1, lowres_cond=True, deep_feature=False)
    ropt = torch.optim.Adam([v for v in sdg.values() if v.requires_grad], lr=1e-6, fused=True)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for name, amp in (("torch_autograd_fp32_tf32_ms", False), ("torch_autograd_bf16_autocast_ms", True)):
            def ref():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    o = unet_forward(sdg, spec, x, t, lowres_cond_img=lr)
                F.mse_loss(o.float(), target).backward()
                ropt.step()
                ropt.zero_grad(set_to_none=True)
            try:
                ref()
                out[name] = event_ms(ref)
            except Exception as e:  # noqa: BLE001
                out[name] = None
                out[name + "_error"] = f"{type(e).__name__}: {e}"[:200]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    ours = out.get("ms_graph_replay") or out["ms_launched_from_python"]
    refs = [v for k, v in out.items() if k.startswith("torch_autograd") and isinstance(v, float)]
    if refs:
        out["speedup_over_fastest_torch_mode"] = min(refs) / ours
    out["workload"] = "one optimizer step, driver U-Net (dim 64), %d^3 patch, batch %d, bf16, l2 loss, Adam" % (size, batch)
    del sdg, unet
    torch.cuda.empty_cache()
    return out


def volume_record(dev, rank, world, dist, side=256, timesteps=20, batch=None):
    """BASELINE config 3 through the same process group: one synthetic `side`^3 low-field volume cut into overlapping 64^3 patches
    (stride 32: 7^3 = 343 for 256^3, data.py:159-162), 5 % skip rule, contiguous shards over the ranks (343 = 8 * 43 - 1: the last rank
    is one patch short and padded), T denoising iterations per patch (eval_config.yaml:21), ONE all-gather, device-side stitch and
    background mask (test_all.py:182-300).  Strong scaling: the volume is fixed, ranks split it."""
    from diffusioniqt_b200 import volume as V
    from diffusioniqt_b200.synth import synthetic_field
    imagen = build_model(timesteps, dev, "bf16")
    imagen.return_host_lists = False
    low = synthetic_field((side, side, side), 11)
    low[: side // 8] = low.min()                                  # air: some patches fall under the 5 % rule
    low = low.to(dev)
    raw = low - low.min()

    def sample_fn(lr):
        return imagen.sample(batch_size=lr.shape[0], start_image_or_video=lr, start_at_unet_number=2, use_tqdm=False)[0]

    def run(vol, rawv):
        return V.infer_volume(sample_fn, vol, patch=64, overlap=32, raw_lowres=rawv, batch_size=batch, fill_value=MIN_BOUND, rank=rank, world=world)

    # warm-up: engine build + graph capture for every batch size the shards will use (full chunks and the ragged tail)
    grid = V.patch_grid(low.shape, 64, 32)
    kept = [o for o in grid if V.keep_patch(raw, o, 64)]
    start, stop, per = V.shard_range(len(kept), rank, world)
    if batch is None:
        # the shard in equal sampler calls of at most 49 patches (343 = 7 x 49 on one GPU, 43 per call from two GPUs up): the step runs at
        # 0.66 of the sustained peak with 7 patches per call and at 0.72 with 32 (profiles/bench_configs_r2j.jsonl); ~20 GB of activations
        calls = max(1, -(-per // 49))
        batch = max(1, -(-per // calls))
    sizes = {min(batch, stop - start - b0) for b0 in range(0, stop - start, batch)}
    lr0 = torch.zeros((1, 1, 64, 64, 64), device=dev)
    for bsz in sorted(sizes):
        sample_fn(lr0.expand(bsz, -1, -1, -1, -1).contiguous())
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = run(low, raw)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    if dist is not None:
        t = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t)
    vol = res.volume
    ok = bool(torch.isfinite(vol).all()) and float(vol.min()) >= min(MIN_BOUND, float(low.min())) - 1e-5
    del imagen
    torch.cuda.empty_cache()
    return dict(workload=f"BASELINE config 3: {side}^3 volume, 64^3 patches, stride 32, T = {timesteps}, {batch} patches per sampler call, bf16",
                seconds=sec, patches=res.n_patches, skipped=res.n_skipped, patches_per_rank=res.patches_per_rank, patches_per_s=res.n_patches / sec,
                n_gpus=world, scaling="strong", finite_and_bounded=ok, gather_bytes=world * res.patches_per_rank * 64 ** 3 * 4,
                timing="CUDA events on rank-local stream around gather -> sample -> all_gather -> stitch, max over ranks")


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    per_step = max(5, args.ref_denoise_steps)         # >= 5 timed denoising iterations per bench step; the spread is reported
    vals, its = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(args.size, args.timesteps, args.batch, per_step, warm=(i == 0))
        if i >= args.warmup:
            vals.append(r["value"])
            its += [r["ms_per_denoise_step_min"], r["ms_per_denoise_step"], r["ms_per_denoise_step_max"]]
    v = statistics.median(vals)
    r["value"] = v
    r["spread"] = dict(value_min=min(vals), value_max=max(vals), ms_per_denoise_step_min=min(its), ms_per_denoise_step_max=max(its),
                       bench_steps=len(vals), iterations_per_bench_step=per_step)
    line = dict(impl="reference", metric="3D patches/sec (full denoise loop)", value=v, unit="patches/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 / v * args.batch, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload_config(args), cpu_baseline=r,
                e2e=dict(value=v, unit="patches/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="reference CPU algorithm (oracle port) on host cores; one bench step = a bounded sample, value extrapolated to the full sampler")
    print(json.dumps(line), flush=True)


def workload_config(args):
    return dict(workload=f"BASELINE config 2: single {args.size}^3 patch, SRUnet256 driver config (dim 64, mults 1-2-4, 2 resnet blocks/level, SE, "
                         f"deep_feature off), {args.timesteps}-step DDPM sampler, batch {args.batch} per GPU",
                patch=args.size, batch_per_gpu=args.batch, timesteps=args.timesteps, parallelism=f"patch-parallel x{args.gpus}",
                l2="per-step working set ~0.4 GB of activations > 126 MB L2; no explicit flush")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-denoise-steps", type=int, default=3, help="bounded CPU-baseline sample (denoising iterations)")
    ap.add_argument("--ref-denoise-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the training-step sub-record (SURVEY 8 f-4)")
    ap.add_argument("--no-volume", action="store_true", help="skip the BASELINE config 3 sub-record (whole 256^3 volume, T = 20)")
    ap.add_argument("--volume-side", type=int, default=256)
    ap.add_argument("--volume-timesteps", type=int, default=20)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    from diffusioniqt_b200 import lib as L
    from diffusioniqt_b200.synth import synthetic_field
    L.load()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # "NCCL version ..." goes to stdout and would precede the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    B, S = args.batch, args.size
    # the isolated-kernel timings run FIRST, on a GPU that has not yet been under minutes of full load: the burst peak they are divided by
    # (MEASURED_PEAKS.json: best of 10 cuBLAS calls) was taken in the same condition
    dom = time_dominant_kernel(S, B) if rank == 0 else None
    ew = time_elementwise_kernel(S, B) if rank == 0 else None
    barrier()
    imagen = build_model(args.timesteps, dev, args.dtype)
    unet, sched = imagen.unets[1], imagen.noise_schedulers[1]
    shape = (B, 1, S, S, S)
    lr_host = synthetic_field(shape, 100 + rank).pin_memory()
    lr_dev = lr_host.to(dev)
    gathered = torch.empty((world,) + shape, dtype=torch.float32, device=dev) if dist is not None else None

    def device_step():
        imagen.return_host_lists = False
        img, _, _ = imagen.p_sample_loop(unet, shape, noise_scheduler=sched, lowres_cond_img=lr_dev, pred_objective="x_start",
                                         dynamic_threshold=False, use_tqdm=False)
        if dist is not None:
            dist.all_gather_into_tensor(gathered, img)       # the only collective on the path: denoised patches for stitching
        return img

    def e2e_step():
        imagen.return_host_lists = True
        img, _, lst = imagen.sample(batch_size=B, start_image_or_video=lr_host, start_at_unet_number=2, use_tqdm=False)
        if dist is not None:
            dist.all_gather_into_tensor(gathered, img)
        return img.cpu()

    torch.manual_seed(1234 + rank)
    for _ in range(args.warmup):
        device_step()
    launches0 = L.launch_count()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = device_step()
    e1.record()
    barrier()
    clk = clocks.stop()
    elapsed_ms = e0.elapsed_time(e1)
    eager_launches = L.launch_count() - launches0
    graph_launches = imagen.last_graph_launches * args.timesteps * args.steps
    assert torch.isfinite(out).all(), "non-finite sampler output"

    # ---- end to end through the public API with host buffers
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    if dist is not None:
        t = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = t.tolist()

    volume = None
    if not args.no_volume:
        volume = volume_record(dev, rank, world, dist, side=args.volume_side, timesteps=args.volume_timesteps)

    if rank == 0:
        peaks = read_peaks()
        patches = world * B * args.steps
        value = patches / (elapsed_ms * 1e-3)
        ms_per_step = elapsed_ms / args.steps
        flops_per_patch = FLOPS_PER_FWD_64 * (S / 64.0) ** 3 * args.timesteps
        step_tflops = value / world * flops_per_patch / 1e12
        traffic, traffic_note = read_traffic("%s 64^3" % dom["kernel"]) if (S == 64 and B == 1) else (None, "only captured for the 64^3 batch-1 shape")
        roofline = dict(bound="tensor", kernel="%s 3x3x3 64->64 @%d^3 (batch %d), %s" % (dom["kernel"], S, B, dom["as_launched"]), achieved=dom["tflops"],
                        peak=peaks["burst"], unit="TFLOP/s", frac=dom["tflops"] / peaks["burst"], traffic=traffic,
                        plain_conv=dict(achieved=dom["plain"]["tflops"], frac=dom["plain"]["tflops"] / peaks["burst"], ms_per_launch=dom["plain"]["ms"],
                                        note="the same conv launched without the fused input normalisation (what the r1 line measured); in the step "
                                             "that variant needs a separate GroupNorm-apply kernel in front of it"),
                        timing="20 launches over 4 rotating buffer sets captured as one CUDA graph (back to back with programmatic dependent launch, "
                               "as inside a sampler iteration), CUDA events around 3 replays",
                        traffic_note=traffic_note + "; algorithmic traffic is 67.1 MB (32 MiB in + 32 MiB out), the output of a launch stays in the 126 MB L2",
                        ms_per_launch=dom["ms"], flops_per_launch=dom["flops"],
                        peak_source=peaks["source"] + " (burst: kernel timed alone)",
                        whole_step=dict(achieved=step_tflops, peak=peaks["sustained"], frac=step_tflops / peaks["sustained"], unit="TFLOP/s",
                                        note="all FLOPs of the U-Net / wall time of the sampler, per GPU, vs sustained bf16 peak"))
        roofline_hbm = dict(bound="hbm", kernel="scale_residual_ring_kernel (SE gate + h*gate + residual + statistics) 64 ch @%d^3 (batch %d)" % (S, B),
                            achieved=ew["gbs"], peak=peaks["hbm"], unit="GB/s", frac=ew["gbs"] / peaks["hbm"], ms_per_launch=ew["ms"],
                            bytes_per_launch=ew["bytes"],
                            note="second-largest kernel class of the step (19 launches per iteration; the GroupNorm apply now rides on the convs); "
                                 "algorithmic bytes = 2 reads + 1 write, 20 launches over rotating buffers replayed as one CUDA graph, CUDA events around "
                                 "the replays")
        line = dict(metric="3D patches/sec (full denoise loop)", value=value, unit="patches/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic",
                    config=workload_config(args), clocks=clk,
                    e2e=dict(value=patches / (e2e_ms * 1e-3), unit="patches/s", h2d_bytes_per_step=B * S ** 3 * 4, d2h_bytes_per_step=3 * B * S ** 3 * 4),
                    gpu_launches=int(graph_launches + eager_launches), ms_per_denoise_iteration=ms_per_step / args.timesteps,
                    roofline=roofline, roofline_elementwise=roofline_hbm)
        if volume is not None:
            line["volume"] = volume
        if not args.no_torch_gpu_baseline:
            line["torch_gpu_baseline"] = torch_gpu_baseline(S, B)
            ours = ms_per_step / args.timesteps
            for k, v in line["torch_gpu_baseline"].items():
                if isinstance(v, dict) and "ms_per_denoise_iteration" in v:
                    v["speedup_of_this_library"] = v["ms_per_denoise_iteration"] / ours
        if not args.no_train_step:
            try:
                line["train_step"] = train_step_record(S, B)
            except Exception as e:  # a failure of the secondary record must not take the headline line with it
                line["train_step"] = dict(error=f"{type(e).__name__}: {e}"[:300])
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(S, args.timesteps, B, args.cpu_denoise_steps)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
