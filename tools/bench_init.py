"""init_conv_tc alone at the benchmark shape (two fp32 input planes of 64^3 -> 64 bf16 channels + statistics): CUDA events around back-to-back
launches.  A/B against another build with DIQT_LIB_PATH.    python tools/bench_init.py [size] [c_out]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusioniqt_b200 import lib as L

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c_out = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lib = L.load()
n, c_in, dims = 1, 2, (size,) * 3
vox = size ** 3
x = torch.randn(n, c_in, *dims, device="cuda")
w2 = torch.zeros(c_out, 64)
w2[:, :27 * c_in] = torch.randn(c_out, 27 * c_in) * (27 * c_in) ** -0.5
rows, chunks = torch.arange(c_out)[:, None], torch.arange(8)[None, :]
wsw = torch.gather(w2.bfloat16().view(c_out, 8, 8), 1, (chunks ^ (rows & 7))[:, :, None].expand(c_out, 8, 8)).contiguous().cuda()
bias = torch.randn(c_out, device="cuda") * 0.1
planes = (C.c_void_p * c_in)(*[x.data_ptr() + ci * vox * 4 for ci in range(c_in)])
strides = (C.c_int64 * c_in)(*[c_in * vox] * c_in)
nb, ng = C.c_int(0), C.c_int(0)
L.check(lib.diqt_init_conv_tc_blocks(n, *dims[:2], C.byref(nb)))
L.check(lib.diqt_stats_groups(nb.value, 1, C.byref(ng)))
outs = [torch.empty(n, *dims, c_out, dtype=torch.bfloat16, device="cuda") for _ in range(4)]   # 4 x 32 MiB: rotating outputs
part = torch.zeros(n, nb.value, c_out, 2, device="cuda")
grp = torch.zeros(n, max(ng.value, 1), c_out, 2, device="cuda")
tick = torch.zeros(16 * n, dtype=torch.int32, device="cuda")
st = L.current_stream()
def run(i):
    L.check(lib.diqt_init_conv_tc(planes, strides, c_in, wsw.data_ptr(), bias.data_ptr(), outs[i % 4].data_ptr(), c_out, n, *dims, c_out, part.data_ptr(),
                                  grp.data_ptr(), tick.data_ptr(), st), "init_conv_tc")
for i in range(8):
    run(i)
torch.cuda.synchronize()
reps = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    run(i)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
print("init_conv_tc %d^3 -> %d ch: %.2f us per launch (%d CTAs), %.0f GB/s of output + input" % (size, c_out, us, nb.value, (vox * (c_out * 2 + c_in * 4)) / us / 1e3))
