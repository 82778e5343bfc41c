"""Per-op device time of one U-Net forward in situ (warm L2, real predecessor / successor), CUDA events around every engine op.
Events between launches serialise the PDL overlap, so the sum is a little above the graph replay time; shares are what matter.
    python tools/op_timing.py [size] [reps]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusioniqt_b200 import lib as L

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")
imagen = bench.build_model(4, dev, "bf16")
unet = imagen.unets[1]
eng = unet.engine_for(1, (size,) * 3, dev)
eng.load_inputs(torch.randn(1, 1, size, size, size, device=dev), torch.randn(1, 1, size, size, size, device=dev))
eng.set_condition(torch.zeros(1, device=dev))
eng.film_row_ptr, eng.film_stride_n = 0, 1
names = []
orig_check = L.check
def spy(rc, what=""):
    names.append(what)
    return orig_check(rc, what)
import diffusioniqt_b200.engine as E
E.L.check = spy
st = L.current_stream()
for op in eng._ops:          # discover op names (one check() per op)
    n0 = len(names); op(st); names[n0:] = [names[n0] if len(names) > n0 else "?"]
E.L.check = orig_check
torch.cuda.synchronize()
nops = len(eng._ops)
acc = [0.0] * nops
for r in range(reps + 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(nops + 1)]
    ev[0].record()
    for i, op in enumerate(eng._ops):
        op(st)
        ev[i + 1].record()
    torch.cuda.synchronize()
    if r >= 2:
        for i in range(nops):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) * 1e3 / reps
kind = collections.defaultdict(lambda: [0, 0.0])
for n, t in zip(names, acc):
    k = n.split(".")[-1] if "." in n else n
    if "project" in n or "res_conv" in n or n.startswith("downs") and n.endswith(("4", "4.1")) or "net.0" in n:
        k = "conv:" + k
    kind[k][0] += 1; kind[k][1] += t
tot = sum(acc)
print(f"size {size}^3: {nops} ops, sum {tot:.1f} us")
for k, (c, t) in sorted(kind.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} x{c:3d} {t:9.1f} us {100 * t / tot:5.1f}%  avg {t / c:7.2f}")
print("first 40 ops:")
for n, t in list(zip(names, acc))[:40]:
    print(f"  {n:44s} {t:8.2f}")
