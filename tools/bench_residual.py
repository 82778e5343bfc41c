"""scale_residual_kernel (SE gate + h * gate + residual + statistics of the result) alone: graph replay over rotating buffers,
with the gate from grouped statistics (as in the step), with a precomputed gate, and without statistics output."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusioniqt_b200 import lib as L
lib = L.load()
peaks = bench.read_peaks()
dev = torch.device("cuda")

def graph_time(fn, nb, reps=20):
    for i in range(nb): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn(i % nb)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3

for (S, c) in ((64, 64), (32, 128), (32, 64), (16, 128)):
    vox = S ** 3
    nb = 4
    hs = [torch.randn(vox, c, device=dev).bfloat16() for _ in range(nb)]
    xs = [torch.randn(vox, c, device=dev).bfloat16() for _ in range(nb)]
    os_ = [torch.empty(vox, c, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    hidden = c // 16
    w1, w2 = torch.randn(hidden, c, device=dev) * 0.1, torch.randn(c, hidden, device=dev) * 0.1
    nblk = max(1, min(vox // 128, 148))
    ng = C.c_int(0)
    L.check(lib.diqt_stats_groups(nblk, 1, C.byref(ng)))
    part = torch.zeros(nblk * c * 2, device=dev); grp = torch.zeros(16 * c * 2, device=dev); tick = torch.zeros(16, dtype=torch.int32, device=dev)
    L.check(lib.diqt_channel_stats_g(hs[0].data_ptr(), L.BF16, 1, vox, c, c, nblk, part.data_ptr(), grp.data_ptr(), tick.data_ptr(), L.current_stream()))
    opart = torch.zeros(nblk * c * 2, device=dev); ogrp = torch.zeros(16 * c * 2, device=dev); otick = torch.zeros(16, dtype=torch.int32, device=dev)
    gate = torch.rand(c, device=dev)
    def full(i):
        L.check(lib.diqt_scale_residual_g(hs[i].data_ptr(), c, xs[i].data_ptr(), c, os_[i].data_ptr(), c, L.BF16, 1, vox, c, grp.data_ptr(), ng.value, hidden,
                                          w1.data_ptr(), w2.data_ptr(), nblk, opart.data_ptr(), ogrp.data_ptr(), otick.data_ptr(), L.current_stream()))
    def pre_gate(i):
        L.check(lib.diqt_scale_residual(hs[i].data_ptr(), c, xs[i].data_ptr(), c, os_[i].data_ptr(), c, L.BF16, 1, vox, c, gate.data_ptr(), nblk, opart.data_ptr(), 0, 0, L.current_stream()))
    def no_stats(i):
        L.check(lib.diqt_scale_residual(hs[i].data_ptr(), c, xs[i].data_ptr(), c, os_[i].data_ptr(), c, L.BF16, 1, vox, c, gate.data_ptr(), nblk, 0, 0, 0, L.current_stream()))
    nbytes = 3.0 * vox * c * 2
    r = dict(side=S, channels=c, mbytes=nbytes / 1e6)
    for name, fn in (("gate_from_grouped_stats", full), ("precomputed_gate", pre_gate), ("precomputed_gate_no_stats", no_stats)):
        t = graph_time(fn, nb)
        r[name + "_us"] = t
        r[name + "_frac_hbm"] = nbytes / t / 1e3 / peaks["hbm"]
    print(json.dumps(r), flush=True)
