"""Per-CTA wall-clock (globaltimer) of conv_zm_kernel inside a graph of 20 back-to-back launches (diagnostic build with per-CTA stamps,
DIQT_LIB_PATH): entry, set-up done, grid dependency released (plane producer past griddepcontrol.wait), CTA end -- for the last launches."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from diffusioniqt_b200 import lib as L
kind = sys.argv[1] if len(sys.argv) > 1 else "fused"
out = bench.time_dominant_kernel(64, 1, kinds=(kind,))
raw = C.CDLL(L.LIB_PATH)
buf = (C.c_ulonglong * (8 * 4 * 160))()
assert raw.diqt_debug_zm_cta(buf) == 0
a = np.array(buf, dtype=np.float64).reshape(8, 4, 160)[:, :, :144] / 1000.0     # us
order = np.argsort(a[:, 0].min(axis=1))
a = a[order]
t0 = a[0, 0].min()
print(f"conv_zm 64->64 @64^3 {kind}: {out['ms'] * 1e3:.2f} us per launch (graph of 20); the last 8 launches, per-CTA stamps in us from the first entry")
prev_end = None
for k in range(8):
    e0, e1, e2, e3 = (a[k, i] - t0 for i in range(4))
    line = (f"launch {k}: entry {e0.min():8.2f} .. {e0.max():8.2f} | set-up {np.median(e1 - e0):5.2f} | dependency released {e2.min():8.2f} .. {e2.max():8.2f} | "
            f"end {e3.min():8.2f} .. {e3.max():8.2f} (median {np.median(e3):8.2f}) | CTA busy (release -> end) median {np.median(e3 - e2):6.2f} max {(e3 - e2).max():6.2f}")
    if prev_end is not None:
        line += f" | previous grid's last CTA end -> first release {e2.min() - prev_end:5.2f}, period {e3.max() - prev_end:6.2f}"
    prev_end = e3.max()
    print(line)
busy = (a[6, 3] - a[6, 2])
print("CTA busy (us) by block index, launch 6, rows of 16:")
for r in range(0, 144, 16):
    print("  " + " ".join(f"{v:5.1f}" for v in busy[r:r + 16]))
