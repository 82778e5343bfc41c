"""Softmax attention kernels alone (BASELINE config 5: N in {1728, 13824} tokens, 8 heads x 64): tcgen05 kernel vs the CUDA-core kernel.
FLOPs counted as 4 N^2 d per head (Q K^T and P V once each); the two-pass tensor-core kernel executes 6 N^2 d."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
from diffusioniqt_b200 import lib as L
import bench

lib = L.load()
peaks = bench.read_peaks()
heads, dh = 8, 64
inner = heads * dh
for n in (1728, 13824):
    qkv = (torch.randn(n, 3 * inner, device="cuda")).bfloat16()
    out = torch.empty(n, inner, dtype=torch.bfloat16, device="cuda")
    p, esz = qkv.data_ptr(), 2
    nbytes = C.c_size_t(0)
    L.check(lib.diqt_attn_tc_workspace_bytes(n, heads, C.byref(nbytes)))
    ws = torch.zeros(nbytes.value, dtype=torch.uint8, device="cuda")
    plan = C.c_void_p(0)
    L.check(lib.diqt_attn_tc_plan_create(p, p + inner * esz, p + 2 * inner * esz, 3 * inner, 3 * inner, 3 * inner, out.data_ptr(), inner, n, heads, dh ** -0.5, 1,
                                         ws.data_ptr(), C.byref(plan)))
    st = L.current_stream()

    def run_tc():
        L.check(lib.diqt_attn_tc_run(plan.value, st))

    def run_simt():
        L.check(lib.diqt_softmax_attention(p, p + inner * esz, p + 2 * inner * esz, 3 * inner, 3 * inner, 3 * inner, out.data_ptr(), inner, L.BF16, n, heads, dh,
                                           dh ** -0.5, 1, st))

    for name, fn, reps in (("tcgen05", run_tc, 10), ("cuda-core", run_simt, 2 if n > 2000 else 10)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 4.0 * n * n * dh * heads
        print(json.dumps(dict(kernel=name, tokens=n, heads=heads, dim_head=dh, ms=ms, tflops_algorithmic=fl / ms / 1e9,
                              frac_burst_peak=fl / ms / 1e9 / peaks["burst"])), flush=True)
    lib.diqt_attn_tc_plan_destroy(plan.value)
