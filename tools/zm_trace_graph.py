"""In-graph timeline of conv_zm_kernel (build variant -DDIQT_ZM_TRACE=1 via DIQT_LIB_PATH): bench.time_dominant_kernel's 20 back-to-back
launches with programmatic dependent launch, trace of CTA 0 of the LAST launch (times from its first event, microseconds at 1965 MHz)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from diffusioniqt_b200 import lib as L
kind = sys.argv[1] if len(sys.argv) > 1 else "fused"
out = bench.time_dominant_kernel(64, 1, kinds=(kind,))
raw = C.CDLL(L.LIB_PATH)
buf = (C.c_longlong * (8 * 2 * 16))()
assert raw.diqt_debug_zm_trace(buf) == 0
names = ["TMA issue", "plane landed (xf start)", "xf done", "issuer has plane", "issuer issued plane's MMAs", "epilogue: acc complete", "epilogue: drained"]
t0 = min(v for v in buf if v > 0)
print(f"conv_zm 64->64 @64^3 {kind}, in a graph of 20 back-to-back launches: {out['ms'] * 1e3:.2f} us per launch; CTA 0 of the last launch")
for s in range(2):
    print(f"slot {s}:  iter " + " ".join(f"{i:7d}" for i in range(10)))
    for e, nm in enumerate(names):
        row = [buf[(e * 2 + s) * 16 + i] for i in range(10)]
        print(f"  {nm:28s}" + " ".join(f"{(v - t0) / 1965.0:7.2f}" if v > 0 else "      -" for v in row))

# tail events (diagnostic builds that record them): epilogue after its last plane, statistics flushed, group tail done, stores drained; CTA-wide end barrier
ev7 = [buf[(7 * 2 + 0) * 16 + i] for i in range(4)] + [buf[(7 * 2 + 1) * 16 + i] for i in range(2)]
if any(ev7):
    labels = ["epilogue loop done", "statistics flushed", "group tail done", "stores drained", "thread 0 at the end barrier", "end barrier passed"]
    print("tail events: " + "; ".join(f"{l} {(v - t0) / 1965.0:.2f}" for l, v in zip(labels, ev7) if v > 0))
