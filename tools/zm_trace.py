"""Timeline of one CTA of conv_zm_kernel (build variant -DDIQT_ZM_TRACE=1, DIQT_LIB_PATH pointing at it): clock64 at the hand-over points
of the plane pipeline, relative to the first event, in microseconds at the nominal SM clock."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusioniqt_b200 import lib as L, ops
import torch.nn.functional as F
fused = (sys.argv[1] if len(sys.argv) > 1 else "fused") == "fused"
pair = (sys.argv[2] if len(sys.argv) > 2 else "pair") == "pair"
lib = L.load()
raw = C.CDLL(L.LIB_PATH)
S, c = 64, 64
x = torch.randn(1, S, S, S, c, device="cuda").bfloat16()
w = torch.randn(c, c, 3, 3, 3) * 0.02
b = torch.zeros(c)
gn = dict(groups=8, gamma=torch.ones(c), beta=torch.zeros(c), scale_shift=torch.zeros(1, 2 * c), nblk=148) if fused else None
for _ in range(3):
    ops.conv3d(x, w, b, mode="k3", impl="zm", with_stats=True, grouped=True, gn=gn, pair=pair)
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 2 * 16))()
assert raw.diqt_debug_zm_trace(buf) == 0
names = ["TMA issue", "plane landed (xf start)", "xf done", "issuer has plane", "issuer issued plane's MMAs", "epilogue: acc complete", "epilogue: drained"]
t0 = min(v for v in buf if v > 0)
MHZ = 1965.0
print(f"conv_zm 64->64 @64^3, {'fused GN' if fused else 'plain'}, {'CTA pair' if pair else 'single CTA'}; CTA 0, times in us from its first event")
for s in range(2):
    print(f"slot {s}:  iter " + " ".join(f"{i:7d}" for i in range(10)))
    for e, nm in enumerate(names):
        row = [buf[(e * 2 + s) * 16 + i] for i in range(10)]
        print(f"  {nm:28s} " + " ".join(f"{(v - t0) / MHZ:7.2f}" if v > 0 else "      -" for v in row))
