"""Per-kernel parity of the attention-block kernels (csrc/attn.cu) against the PyTorch ops the reference calls
(imagen_pytorch3D.py:361-382, 858-869, 913-924, 986-1016, 1078-1106, 1108-1116), on the CPU in fp32.

Tolerances as in test_gpu_kernels.py: 1e-5 relative in fp32 mode, 1e-2 in bf16 mode (max |a-b| / max |b|)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_rel
from oracle.attn_oracle import chan_layernorm as ln_oracle
from oracle.unet_oracle import merge_sub_volumes

pytestmark = pytest.mark.gpu

DTYPES = [("fp32", torch.float32, 1e-5), ("bf16", torch.bfloat16, 1e-2)]


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _q(x, dt):
    """Round to the activation dtype (so both sides see the same inputs)."""
    return x.to(dt).float()


def _rows(v):
    """(1, c, d, h, w) -> (d*h*w, c)"""
    return v[0].permute(1, 2, 3, 0).reshape(-1, v.shape[1]).contiguous()


def _native_rows(sub):
    """(b, c, a, a, a) -> (b*a^3, c): the engine's sub-volume order"""
    return sub.permute(0, 2, 3, 4, 1).reshape(-1, sub.shape[1]).contiguous()


@pytest.mark.parametrize("name,dt,tol", DTYPES)
@pytest.mark.parametrize("c", [32, 64, 200])
def test_chan_layernorm(name, dt, tol, c):
    from diffusioniqt_b200 import ops
    x = _q(_rand(1, c, 5, 6, 7, seed=1, scale=2.0) + 0.5, dt)
    g = _rand(c, 1, 1, 1, seed=2)
    want = _rows(ln_oracle(x, g))
    got = ops.chan_layernorm(_rows(x).to(dt).cuda(), g)
    assert max_rel(got.float().cpu(), want) < tol
    # GELU in front (ChanFeedForward :1112-1114) and two residuals behind
    r1, r2 = _q(_rand(5 * 6 * 7, c, seed=3), dt), _q(_rand(5 * 6 * 7, c, seed=4), dt)
    want = _rows(ln_oracle(F.gelu(x), g)) + r1 + r2
    got = ops.chan_layernorm(_rows(x).to(dt).cuda(), g, pre_act=2, res1=r1.to(dt).cuda(), res2=r2.to(dt).cuda())
    assert max_rel(got.float().cpu(), want) < tol
    # nn.LayerNorm with bias (ViT :725)
    beta = _rand(c, seed=5)
    want = F.layer_norm(_rows(x), (c,), g.reshape(-1), beta)
    got = ops.chan_layernorm(_rows(x).to(dt).cuda(), g, beta=beta)
    assert max_rel(got.float().cpu(), want) < tol


@pytest.mark.parametrize("name,dt,tol", DTYPES)
def test_chan_layernorm_merged_source(name, dt, tol):
    """x stored merged, out / residual in sub-volume order (the last norm of an attention block)."""
    from diffusioniqt_b200 import ops
    f, h, c = 3, 4, 32
    sub = _q(_rand(f ** 3, c, h, h, h, seed=6), dt)
    res = _q(_rand(f ** 3, c, h, h, h, seed=7), dt)
    g = _rand(c, 1, 1, 1, seed=8)
    want = _native_rows(ln_oracle(sub, g) + res)
    merged = _rows(merge_sub_volumes(sub, f))
    got = ops.chan_layernorm(merged.to(dt).cuda(), g, res1=_native_rows(res).to(dt).cuda(), x_sub=(f, h))
    assert max_rel(got.float().cpu(), want) < tol


@pytest.mark.parametrize("name,dt,tol", DTYPES)
def test_rows_combine(name, dt, tol):
    from diffusioniqt_b200 import ops
    a, b, c2 = (_q(_rand(77, 48, seed=s), dt) for s in (1, 2, 3))
    for act, fn in ((0, lambda t: t), (1, F.mish), (2, F.gelu)):
        got = ops.rows_combine(a.to(dt).cuda(), act, b.to(dt).cuda(), c2.to(dt).cuda())
        assert max_rel(got.float().cpu(), fn(a) + b + c2) < tol
    got = ops.rows_combine(a.to(dt).cuda(), 1)
    assert max_rel(got.float().cpu(), F.mish(a)) < tol


@pytest.mark.parametrize("name,dt,tol", DTYPES)
@pytest.mark.parametrize("f,h,p,c", [(3, 8, 8, 32), (3, 4, 2, 64), (2, 8, 4, 40), (1, 8, 4, 32)])
def test_dw_patchify(name, dt, tol, f, h, p, c):
    from diffusioniqt_b200 import ops
    sub = _q(_rand(f ** 3, c, h, h, h, seed=1), dt)
    w, b = _rand(c, 1, p, p, p, seed=2, scale=p ** -1.5), _rand(c, seed=3, scale=0.1)
    merged = merge_sub_volumes(sub, f)
    want = _rows(F.conv3d(merged, w, b, stride=p, groups=c))
    g = f * h // p
    got = ops.dw_patchify(_native_rows(sub).to(dt).cuda(), w, b, g, p, x_sub=(f, h) if f > 1 else (0, 0))
    assert max_rel(got.float().cpu(), want) < tol
    got = ops.dw_patchify(_rows(merged).to(dt).cuda(), w, b, g, p)        # already merged (boundary mode)
    assert max_rel(got.float().cpu(), want) < tol


@pytest.mark.parametrize("name,dt,tol", DTYPES)
@pytest.mark.parametrize("dims,c,bias", [((3, 3, 3), 96, False), ((6, 5, 7), 32, True), ((12, 12, 12), 64, True)])
def test_dw_conv3(name, dt, tol, dims, c, bias):
    from diffusioniqt_b200 import ops
    x = _q(_rand(1, c, *dims, seed=1), dt)
    w = _rand(c, 1, 3, 3, 3, seed=2, scale=0.2)
    b = _rand(c, seed=3, scale=0.1) if bias else None
    want = F.conv3d(x, w, b, padding=1, groups=c)[0].permute(1, 2, 3, 0)
    got = ops.dw_conv3(x[0].permute(1, 2, 3, 0).contiguous().to(dt).cuda(), w, b)
    assert max_rel(got.float().cpu(), want) < tol


@pytest.mark.parametrize("name,dt,tol", DTYPES)
@pytest.mark.parametrize("g,p,c", [(3, 8, 32), (3, 2, 64), (12, 2, 16), (1, 4, 32)])
def test_upsample_trilinear(name, dt, tol, g, p, c):
    from diffusioniqt_b200 import ops
    x = _q(_rand(1, c, g, g, g, seed=1), dt)
    want = _rows(F.interpolate(x, scale_factor=p, mode="trilinear", align_corners=True))
    got = ops.upsample_trilinear(_rows(x).to(dt).cuda(), g, p)
    assert max_rel(got.float().cpu(), want) < tol


def _split_heads(t, heads):
    n, inner = t.shape
    return t.reshape(n, heads, inner // heads).permute(1, 0, 2)        # (h, n, d)


@pytest.mark.parametrize("name,dt,tol", DTYPES)
@pytest.mark.parametrize("n,heads,dh", [(27, 2, 16), (216, 4, 32), (1728, 8, 64), (100, 2, 64)])
def test_linear_attention(name, dt, tol, n, heads, dh):
    from diffusioniqt_b200 import ops
    qkv = _q(_rand(n, 3 * heads * dh, seed=1, scale=1.5), dt)
    q, k, v = (_split_heads(t, heads) for t in qkv.chunk(3, dim=1))
    q = q.softmax(dim=-1) * dh ** -0.5
    k = k.softmax(dim=-2)
    out = torch.einsum("bnd,bde->bne", q, torch.einsum("bnd,bne->bde", k, v))
    want = F.mish(out.permute(1, 0, 2).reshape(n, heads * dh))
    got = ops.linear_attention(qkv.to(dt).cuda(), heads, dh, impl="simt")
    assert max_rel(got.float().cpu(), want) < tol


def _linear_attention_want(qkv, heads, dh, act):
    n = qkv.shape[0]
    q, k, v = (_split_heads(t, heads) for t in qkv.chunk(3, dim=1))
    q = q.softmax(dim=-1) * dh ** -0.5
    k = k.softmax(dim=-2)
    out = torch.einsum("bnd,bde->bne", q, torch.einsum("bnd,bne->bde", k, v)).permute(1, 0, 2).reshape(n, heads * dh)
    return F.mish(out) if act else out


@pytest.mark.parametrize("n,heads,act", [(128, 2, 0), (100, 2, 1), (27, 4, 1), (1728, 8, 1), (520, 6, 0), (5000, 4, 1), (13824, 8, 1)])
def test_linear_attention_tensor_core_kernel(n, heads, act):
    """csrc/linattn_tc.cu (tcgen05 k^T v with MN-major operands, tcgen05 q ctx) against the fp32 PyTorch product (imagen_pytorch3D.py:1001-1011);
    bf16, dim_head 64.  Token counts that are not multiples of the 128-token tile exercise the row mask (rows past the end must not enter the
    column sums); 13 824 x 8 is BASELINE config 5's long sequence (37 chunks x 4 head pairs); a second pass uses keys with a large
    dynamic range and a trend along the sequence, so the column maximum comes from far-away chunks."""
    from diffusioniqt_b200 import ops
    dh, dt, tol = 64, torch.bfloat16, 1e-2
    qkv = _q(_rand(n, 3 * heads * dh, seed=4, scale=1.5), dt)
    want = _linear_attention_want(qkv, heads, dh, act)
    got = ops.linear_attention(qkv.to(dt).cuda(), heads, dh, act=act, impl="tc")
    assert torch.isfinite(got).all()
    assert max_rel(got.float().cpu(), want) < tol
    assert torch.equal(got, ops.linear_attention(qkv.to(dt).cuda(), heads, dh, act=act, impl="tc"))      # fixed summation order
    simt = ops.linear_attention(qkv.to(dt).cuda(), heads, dh, act=act, impl="simt")
    assert max_rel(got.float().cpu(), simt.float().cpu()) < tol
    inner = heads * dh
    qkv2 = qkv.clone()
    qkv2[:, inner: 2 * inner] *= torch.linspace(0.3, 6.0, n)[:, None]     # k: the largest entries sit at the end of the sequence
    qkv2[:, :inner] *= 4.0                                                # q: sharp softmax over the head dimension
    qkv2 = _q(qkv2, dt)
    want = _linear_attention_want(qkv2, heads, dh, act)
    got = ops.linear_attention(qkv2.to(dt).cuda(), heads, dh, act=act, impl="tc")
    assert torch.isfinite(got).all()
    assert max_rel(got.float().cpu(), want) < tol


def test_linear_attention_tensor_core_rejects_unsupported_shapes():
    from diffusioniqt_b200 import lib as L, ops
    qkv = torch.zeros(64, 3 * 3 * 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(L.DiqtError):
        ops.linear_attention(qkv, 3, 64, impl="tc")           # odd number of heads
    with pytest.raises(L.DiqtError):
        ops.linear_attention(qkv.float(), 3, 64, impl="tc")   # fp32 exact mode stays on the CUDA cores


@pytest.mark.parametrize("name,dt,tol", DTYPES)
@pytest.mark.parametrize("n,heads,dh,act", [(27, 2, 16, 1), (216, 4, 32, 0), (1728, 8, 64, 1), (100, 2, 64, 1)])
def test_softmax_attention(name, dt, tol, n, heads, dh, act):
    from diffusioniqt_b200 import ops
    qkv = _q(_rand(n, 3 * heads * dh, seed=2), dt)
    q, k, v = (_split_heads(t, heads) for t in qkv.chunk(3, dim=1))
    att = (torch.einsum("bqd,bkd->bqk", q, k) * dh ** -0.5).softmax(dim=-1)
    out = torch.einsum("bnd,bde->bne", att, v).permute(1, 0, 2).reshape(n, heads * dh)
    want = F.mish(out) if act else out
    got = ops.softmax_attention(qkv.to(dt).cuda(), heads, dh, act=act)
    assert max_rel(got.float().cpu(), want) < tol


@pytest.mark.parametrize("version", ["2", "1"])
@pytest.mark.parametrize("n,heads,act", [(128, 1, 0), (100, 2, 1), (300, 3, 0), (1728, 8, 1), (520, 2, 1), (4736, 8, 0), (5000, 5, 1)])
def test_softmax_attention_tensor_core_kernel(n, heads, act, version, monkeypatch):
    """csrc/attn_tc.cu (tcgen05 Q K^T and P V) against the fp32 PyTorch product; bf16, dim_head 64; token counts that are not multiples
    of the 128-key tile exercise the key mask and the query-row guard.  Version 2 = single pass, online softmax with lazy rescaling,
    P through tensor memory (one or two query tiles per CTA: 4736 x 8 and 5000 x 5 take the two-tile path); version 1 = two passes."""
    from diffusioniqt_b200 import ops
    monkeypatch.setenv("DIQT_ATTN_TC_VERSION", version)
    dh, dt, tol = 64, torch.bfloat16, 1e-2
    qkv = _q(_rand(n, 3 * heads * dh, seed=3), dt)
    q, k, v = (_split_heads(t, heads) for t in qkv.chunk(3, dim=1))
    att = (torch.einsum("bqd,bkd->bqk", q, k) * dh ** -0.5).softmax(dim=-1)
    out = torch.einsum("bnd,bde->bne", att, v).permute(1, 0, 2).reshape(n, heads * dh)
    want = F.mish(out) if act else out
    got = ops.softmax_attention(qkv.to(dt).cuda(), heads, dh, act=act, impl="tc")
    assert torch.isfinite(got).all()
    assert max_rel(got.float().cpu(), want) < tol
    # sharper logits (larger scale of q): the row maximum matters
    qkv2 = qkv.clone()
    qkv2[:, : heads * dh] *= 6.0
    qkv2 = _q(qkv2, dt)
    q, k, v = (_split_heads(t, heads) for t in qkv2.chunk(3, dim=1))
    att = (torch.einsum("bqd,bkd->bqk", q, k) * dh ** -0.5).softmax(dim=-1)
    out = torch.einsum("bnd,bde->bne", att, v).permute(1, 0, 2).reshape(n, heads * dh)
    want = F.mish(out) if act else out
    got = ops.softmax_attention(qkv2.to(dt).cuda(), heads, dh, act=act, impl="tc")
    assert max_rel(got.float().cpu(), want) < tol


def test_softmax_attention_tensor_core_13824_tokens():
    """BASELINE config 5's long sequence (192^3 / 8^3 tokens): two heads against fp32 PyTorch on the CPU, plus rows whose maximum keeps
    growing from key tile to key tile (every lazy rescale of O in tensor memory is exercised) and rows with one dominant late key."""
    from diffusioniqt_b200 import ops
    n, heads, dh, dt = 13824, 2, 64, torch.bfloat16
    qkv = _q(_rand(n, 3 * heads * dh, seed=9), dt)
    # keys whose norm grows with their index: the running maximum of every query row rises across the 108 key tiles
    ramp = torch.linspace(0.2, 3.0, n)[:, None]
    qkv[:, heads * dh: 2 * heads * dh] *= ramp
    qkv[:, : heads * dh] *= 3.0
    qkv = _q(qkv, dt)
    q, k, v = (_split_heads(t, heads) for t in qkv.chunk(3, dim=1))
    want = torch.empty(heads, n, dh)
    for h in range(heads):
        for r0 in range(0, n, 1728):
            att = (q[h, r0:r0 + 1728] @ k[h].t() * dh ** -0.5).softmax(dim=-1)
            want[h, r0:r0 + 1728] = att @ v[h]
    want = want.permute(1, 0, 2).reshape(n, heads * dh)
    got = ops.softmax_attention(qkv.to(dt).cuda(), heads, dh, act=0, impl="tc")
    assert torch.isfinite(got).all()
    assert max_rel(got.float().cpu(), want) < 1e-2
