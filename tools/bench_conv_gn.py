"""Block.forward as one launch: conv_zm_kernel with the fused input GroupNorm + FiLM + Mish against the two-kernel path (apply kernel +
plain conv), alone, rotating buffers, graph replay + CUDA events.  Shapes: the U-Net's 3x3x3 convs at the driver config."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusioniqt_b200 import lib as L

lib = L.load()
peaks = bench.read_peaks()
dev = torch.device("cuda")


def graph_time(fns, reps):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fns[i % len(fns)]()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


for (S, ci, co) in ((64, 64, 64), (64, 128, 64), (32, 64, 64), (32, 128, 128), (32, 192, 128), (16, 128, 128)):
    vox = S ** 3
    nb = 4
    xs = [torch.randn(vox, ci, device=dev).bfloat16() for _ in range(nb)]
    As = [torch.empty(vox, ci, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    ys = [torch.empty(vox, co, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    w = (torch.randn(co, ci, 3, 3, 3, device=dev) * (27 * ci) ** -0.5).bfloat16().float().contiguous()
    b = torch.zeros(co, device=dev)
    gamma, beta = torch.rand(ci, device=dev) + 0.5, torch.randn(ci, device=dev) * 0.1
    film = (torch.randn(1, 2 * ci, device=dev) * 0.2).contiguous()
    desc = L.ConvDesc(mode=L.CONV_K3, dtype=L.BF16, impl=L.IMPL_ZM, n=1, d0=S, d1=S, d2=S, c_in=ci, ld_in=ci, c_out=co, ld_out=co, flags=0)
    nbytes = C.c_size_t(0)
    L.check(lib.diqt_conv_packed_bytes(C.byref(desc), C.byref(nbytes)))
    packed = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    pb = torch.empty(co, dtype=torch.float32, device=dev)
    st = L.current_stream()
    L.check(lib.diqt_conv_pack(C.byref(desc), w.data_ptr(), b.data_ptr(), packed.data_ptr(), pb.data_ptr(), st))
    nblk = max(1, min(vox // 128, 148))
    ng = C.c_int(0)
    L.check(lib.diqt_stats_groups(nblk, 1, C.byref(ng)))
    part = torch.zeros(nblk * ci * 2, device=dev)
    grp = torch.zeros(16 * ci * 2, device=dev)
    tick = torch.zeros(16, dtype=torch.int32, device=dev)
    L.check(lib.diqt_channel_stats_g(xs[0].data_ptr(), L.BF16, 1, vox, ci, ci, nblk, part.data_ptr(), grp.data_ptr(), tick.data_ptr(), st))
    opart, ogrp, otick = torch.zeros(320 * co * 2, device=dev), torch.zeros(16 * co * 2, device=dev), torch.zeros(16, dtype=torch.int32, device=dev)
    fused, plain = [], []
    for x, a, y in zip(xs, As, ys):
        for src, lst, gn in ((x, fused, True), (a, plain, False)):
            p = C.c_void_p(0)
            L.check(lib.diqt_conv_plan_create(C.byref(desc), src.data_ptr(), y.data_ptr(), packed.data_ptr(), pb.data_ptr(), C.byref(p)))
            nbk, ngo = C.c_int(0), C.c_int(0)
            L.check(lib.diqt_conv_plan_set_stats_g(p.value, opart.data_ptr(), ogrp.data_ptr(), otick.data_ptr(), C.byref(nbk), C.byref(ngo)))
            if gn:
                L.check(lib.diqt_conv_plan_set_gn(p.value, grp.data_ptr(), ng.value, vox, 8, 1e-5, gamma.data_ptr(), beta.data_ptr()))
                L.check(lib.diqt_conv_plan_set_film(p.value, film.data_ptr(), 2 * ci, 0, 1))
            lst.append(p.value)
    nbk_a = max(1, min(vox // 128, 592))

    def apply(x, a):
        L.check(lib.diqt_gn_mish_g(x.data_ptr(), ci, a.data_ptr(), ci, L.BF16, 1, vox, ci, grp.data_ptr(), ng.value, 8, 1e-5, gamma.data_ptr(), beta.data_ptr(),
                                   film.data_ptr(), 2 * ci, 0, 1, nbk_a, L.current_stream()))

    reps = 20
    t_fused = graph_time([(lambda p=p: L.check(lib.diqt_conv_run(p, L.current_stream()))) for p in fused], reps)
    t_plain = graph_time([(lambda p=p: L.check(lib.diqt_conv_run(p, L.current_stream()))) for p in plain], reps)
    t_two = graph_time([(lambda p=p, x=x, a=a: (apply(x, a), L.check(lib.diqt_conv_run(p, L.current_stream())))) for p, x, a in zip(plain, xs, As)], reps)
    fl = 2.0 * ci * co * 27 * vox
    print(json.dumps(dict(side=S, c_in=ci, c_out=co, gflop=fl / 1e9, fused_us=t_fused, plain_conv_us=t_plain, apply_plus_conv_us=t_two,
                          fused_tflops=fl / t_fused / 1e6, fused_frac_burst=fl / t_fused / 1e6 / peaks["burst"], plain_frac_burst=fl / t_plain / 1e6 / peaks["burst"])), flush=True)
    for p in fused + plain:
        lib.diqt_conv_plan_destroy(p)
