"""One configuration of the tcgen05 softmax attention, a few launches: the target of `ncu -k regex:softmax_attn_tc2_kernel`."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusioniqt_b200 import ops
n, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 13824, int(sys.argv[2]) if len(sys.argv) > 2 else 8
qkv = torch.randn(n, 3 * heads * 64, device="cuda").bfloat16()
for _ in range(3):
    out = ops.softmax_attention(qkv, heads, 64, act=0, impl="tc")
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
