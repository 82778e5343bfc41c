"""Tensor-level wrappers over the C ABI (one call = one or a few kernels).

Used by the per-kernel parity tests and handy for experiments; the engine binds the same entry
points with pre-resolved pointers.  Activations are channels-last: `(n, d0, d1, d2, c)` tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L

_MODES = {"k3": L.CONV_K3, "k1": L.CONV_K1, "down": L.CONV_DOWN, "up": L.CONV_UP}
_IMPLS = {"auto": L.IMPL_AUTO, "simt": L.IMPL_SIMT, "tc": L.IMPL_TC, "zm": L.IMPL_ZM}


SYNC = True   # wrappers wait for their kernels (tests read results right away); the training step switches this off and lets the stream run ahead


def _sync():
    if SYNC:
        torch.cuda.current_stream().synchronize()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return L.BF16
    if t.dtype == torch.float32:
        return L.F32
    raise TypeError(f"unsupported activation dtype {t.dtype}")


def to_channels_last(x: torch.Tensor, dtype=None) -> torch.Tensor:
    """(n, c, d0, d1, d2) -> contiguous (n, d0, d1, d2, c)."""
    y = x.permute(0, 2, 3, 4, 1).contiguous()
    return y.to(dtype) if dtype is not None else y


def from_channels_last(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 4, 1, 2, 3).contiguous().float()


def conv3d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], mode: str = "k3", impl: str = "auto", with_stats: bool = False,
           grouped: bool = False, gn: Optional[dict] = None, pair: bool = True, split_k: bool = True):
    """x: (n, d0, d1, d2, c_in) bf16/fp32 on CUDA; weight/bias exactly as in the reference state_dict (fp32).
    gn = dict(groups, gamma, beta, scale_shift=None, eps=1e-5, nblk=8): the conv computes conv(mish(FiLM(GroupNorm(x)))) with the
    normalisation fused into its load path (z-march family only; raises if the plan cannot).
    pair=False keeps the z-march kernel on single CTAs where it would otherwise run as CTA pairs (tcgen05.mma.cta_group::2).
    split_k=False keeps the per-tap kernel unsplit on small volumes (it otherwise deals a tile's K-blocks to several CTAs).
    with_stats=True also returns the fused per-block channel statistics (n, nblk, c_out, 2) or None if the kernel cannot fuse them;
    with grouped=True the statistics are (partial rows, group sums (n, ngroups, c_out, 2)) through the grouped sink."""
    lib = L.load()
    n, d0, d1, d2, c_in = x.shape
    x = x.contiguous()
    c_out = weight.shape[0]
    if mode == "down":
        assert weight.shape[1] == 8 * c_in
        out = torch.empty(n, d0 // 2, d1 // 2, d2 // 2, c_out, dtype=x.dtype, device=x.device)
        ld_out = c_out
    elif mode == "up":
        out = torch.empty(n, 2 * d0, 2 * d1, 2 * d2, c_out // 8, dtype=x.dtype, device=x.device)
        ld_out = c_out // 8
    else:
        out = torch.empty(n, d0, d1, d2, c_out, dtype=x.dtype, device=x.device)
        ld_out = c_out
    desc = L.ConvDesc(mode=_MODES[mode], dtype=_dt(x), impl=_IMPLS[impl], n=n, d0=d0, d1=d1, d2=d2, c_in=c_in, ld_in=c_in,
                      c_out=c_out, ld_out=ld_out, flags=0 if pair else L.CONV_FLAG_NO_CTA_PAIR)
    nbytes = C.c_size_t(0)
    L.check(lib.diqt_conv_packed_bytes(C.byref(desc), C.byref(nbytes)), "conv_packed_bytes")
    packed = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
    pbias = torch.empty(c_out, dtype=torch.float32, device=x.device)
    w = weight.detach().to(device=x.device, dtype=torch.float32)
    if x.dtype == torch.bfloat16:
        w = w.to(torch.bfloat16).float()
    w = w.contiguous()
    b = bias.detach().to(device=x.device, dtype=torch.float32).contiguous() if bias is not None else None
    st = L.current_stream()
    L.check(lib.diqt_conv_pack(C.byref(desc), w.data_ptr(), L.ptr(b), packed.data_ptr(), pbias.data_ptr(), st), "conv_pack")
    plan = C.c_void_p(0)
    L.check(lib.diqt_conv_plan_create(C.byref(desc), x.data_ptr(), out.data_ptr(), packed.data_ptr(), pbias.data_ptr(), C.byref(plan)), "conv_plan")
    stats = None
    try:
        wsb = C.c_size_t(0)
        L.check(lib.diqt_conv_plan_workspace_bytes(plan.value, C.byref(wsb)), "conv_plan_workspace_bytes")
        if wsb.value and split_k:
            ws = torch.zeros(wsb.value, dtype=torch.uint8, device=x.device)
            L.check(lib.diqt_conv_plan_set_workspace(plan.value, ws.data_ptr(), wsb.value), "conv_plan_set_workspace")
        if gn is not None and gn.get("affine"):
            # any batch / width: statistics -> diqt_gn_finalize -> (a, b) in global memory -> the conv applies mish(a * x + b)
            nblk = gn.get("nblk", 8)
            part = channel_stats(x, nblk)
            vox = d0 * d1 * d2
            ga = gn["gamma"].detach().float().contiguous().to(x.device)
            be = gn["beta"].detach().float().contiguous().to(x.device)
            film = gn.get("scale_shift")
            film = film.detach().float().contiguous().to(x.device) if film is not None else None
            aa = torch.empty(n, c_in, dtype=torch.float32, device=x.device)
            ab = torch.empty_like(aa)
            L.check(lib.diqt_gn_finalize(part.data_ptr(), n, nblk, vox, c_in, gn["groups"], gn.get("eps", 1e-5), ga.data_ptr(), be.data_ptr(), L.ptr(film),
                                         2 * c_in, 0, 1, aa.data_ptr(), ab.data_ptr(), st), "gn_finalize")
            L.check(lib.diqt_conv_plan_set_gn_affine(plan.value, aa.data_ptr(), ab.data_ptr()), "conv_plan_set_gn_affine")
        elif gn is not None:
            if not lib.diqt_conv_gn_fusable(C.byref(desc)):
                raise L.DiqtError("conv3d(gn=...): this shape / family cannot fuse the input GroupNorm")
            _, ggrp = channel_stats_grouped(x, gn.get("nblk", 8))
            gamma = gn["gamma"].detach().float().contiguous().to(x.device)
            beta = gn["beta"].detach().float().contiguous().to(x.device)
            L.check(lib.diqt_conv_plan_set_gn(plan.value, ggrp.data_ptr(), ggrp.shape[1], d0 * d1 * d2, gn["groups"], gn.get("eps", 1e-5),
                                              gamma.data_ptr(), beta.data_ptr()), "conv_plan_set_gn")
            film = gn.get("scale_shift")
            if film is not None:
                film = film.detach().float().contiguous().to(x.device)
                L.check(lib.diqt_conv_plan_set_film(plan.value, film.data_ptr(), 2 * c_in, 0, 1), "conv_plan_set_film")
        if with_stats:
            buf = torch.full((n * 320 * c_out * 2,), float("nan"), dtype=torch.float32, device=x.device)
            nb = C.c_int(0)
            if grouped:
                gbuf = torch.full((n * 16 * c_out * 2,), float("nan"), dtype=torch.float32, device=x.device)
                tick = torch.zeros(16, dtype=torch.int32, device=x.device)
                ng = C.c_int(0)
                L.check(lib.diqt_conv_plan_set_stats_g(plan.value, buf.data_ptr(), gbuf.data_ptr(), tick.data_ptr(), C.byref(nb), C.byref(ng)), "conv_set_stats_g")
                if nb.value > 0:
                    stats = (buf[: n * nb.value * c_out * 2].view(n, nb.value, c_out, 2), gbuf[: n * ng.value * c_out * 2].view(n, ng.value, c_out, 2))
            else:
                L.check(lib.diqt_conv_plan_set_stats(plan.value, buf.data_ptr(), C.byref(nb)), "conv_set_stats")
                if nb.value > 0:
                    stats = buf[: n * nb.value * c_out * 2].view(n, nb.value, c_out, 2)
        L.check(lib.diqt_conv_run(plan.value, st), "conv_run")
        _sync()
    finally:
        lib.diqt_conv_plan_destroy(plan.value)
    return (out, stats) if with_stats else out


def channel_stats(x: torch.Tensor, nblk: int = 8) -> torch.Tensor:
    """-> partial sums (n, nblk, c, 2)."""
    lib = L.load()
    n, c = x.shape[0], x.shape[-1]
    vox = x.numel() // (n * c)
    part = torch.empty(n, nblk, c, 2, dtype=torch.float32, device=x.device)
    L.check(lib.diqt_channel_stats(x.data_ptr(), _dt(x), n, vox, c, c, nblk, part.data_ptr(), 0, 0, L.current_stream()), "channel_stats")
    return part


def channel_stats_grouped(x: torch.Tensor, nblk: int = 8):
    """-> (partial (n, nblk, c, 2), group sums (n, ngroups, c, 2)) through the grouped sink of include/diqt.h."""
    lib = L.load()
    n, c = x.shape[0], x.shape[-1]
    vox = x.numel() // (n * c)
    ng = C.c_int(0)
    L.check(lib.diqt_stats_groups(nblk, 1, C.byref(ng)), "stats_groups")
    part = torch.empty(n, nblk, c, 2, dtype=torch.float32, device=x.device)
    grp = torch.full((n, ng.value, c, 2), float("nan"), dtype=torch.float32, device=x.device)
    tick = torch.zeros(n * ng.value, dtype=torch.int32, device=x.device)
    L.check(lib.diqt_channel_stats_g(x.data_ptr(), _dt(x), n, vox, c, c, nblk, part.data_ptr(), grp.data_ptr(), tick.data_ptr(),
                                     L.current_stream()), "channel_stats_g")
    _sync()
    assert int(tick.abs().sum()) == 0, "tickets must reset themselves"
    return part, grp


def group_norm_film_mish(x: torch.Tensor, groups: int, gamma, beta, scale_shift: Optional[torch.Tensor] = None, eps: float = 1e-5, nblk: int = 8,
                         grouped: bool = False):
    """GroupNorm(groups) -> [x*(scale+1)+shift] -> Mish.  scale_shift: (n, 2c) fp32 rows (scale | shift) or None.
    grouped=True: statistics through the grouped sink and ONE kernel (finalisation in the prologue) instead of finalize + apply."""
    lib = L.load()
    n, c = x.shape[0], x.shape[-1]
    vox = x.numel() // (n * c)
    if grouped:
        _, grp = channel_stats_grouped(x, nblk)
        g = gamma.detach().float().contiguous().to(x.device)
        be = beta.detach().float().contiguous().to(x.device)
        film = scale_shift.detach().float().contiguous().to(x.device) if scale_shift is not None else None
        y = torch.empty_like(x)
        L.check(lib.diqt_gn_mish_g(x.data_ptr(), c, y.data_ptr(), c, _dt(x), n, vox, c, grp.data_ptr(), grp.shape[1], groups, eps, g.data_ptr(),
                                   be.data_ptr(), L.ptr(film), 2 * c, 0, 1, nblk, L.current_stream()), "gn_mish_g")
        _sync()
        return y
    part = channel_stats(x, nblk)
    a = torch.empty(n, c, dtype=torch.float32, device=x.device)
    b = torch.empty_like(a)
    g = gamma.detach().float().contiguous().to(x.device)
    be = beta.detach().float().contiguous().to(x.device)
    film = scale_shift.detach().float().contiguous().to(x.device) if scale_shift is not None else None
    st = L.current_stream()
    L.check(lib.diqt_gn_finalize(part.data_ptr(), n, nblk, vox, c, groups, eps, g.data_ptr(), be.data_ptr(), L.ptr(film), 2 * c, 0, 1,
                                 a.data_ptr(), b.data_ptr(), st), "gn_finalize")
    y = torch.empty_like(x)
    L.check(lib.diqt_affine_mish(x.data_ptr(), c, y.data_ptr(), c, _dt(x), n, vox, c, a.data_ptr(), b.data_ptr(), nblk, 0, 0, st), "affine_mish")
    _sync()
    return y


def se_scale_residual(h: torch.Tensor, res: torch.Tensor, w1, w2, nblk: int = 8, grouped: bool = False, return_stats: bool = False):
    """SE3D gate on h, then h*gate + res.  Returns (out, gate, partial stats of out); grouped=True computes the gate in the
    residual kernel's prologue from grouped statistics and returns (out, None, group sums of out)."""
    lib = L.load()
    n, c = h.shape[0], h.shape[-1]
    vox = h.numel() // (n * c)
    st = L.current_stream()
    if grouped:
        _, grp = channel_stats_grouped(h, nblk)
        w1 = w1.detach().float().contiguous().to(h.device)
        w2 = w2.detach().float().contiguous().to(h.device)
        out = torch.empty_like(h)
        opart = torch.empty(n, nblk, c, 2, dtype=torch.float32, device=h.device)
        ogrp = torch.full((n, grp.shape[1], c, 2), float("nan"), dtype=torch.float32, device=h.device)
        tick = torch.zeros(n * grp.shape[1], dtype=torch.int32, device=h.device)
        L.check(lib.diqt_scale_residual_g(h.data_ptr(), c, res.data_ptr(), c, out.data_ptr(), c, _dt(h), n, vox, c, grp.data_ptr(), grp.shape[1],
                                          w1.shape[0], w1.data_ptr(), w2.data_ptr(), nblk, opart.data_ptr(), ogrp.data_ptr(), tick.data_ptr(), st),
                "scale_residual_g")
        _sync()
        return out, None, ogrp
    part = channel_stats(h, nblk)
    gate = torch.empty(n, c, dtype=torch.float32, device=h.device)
    w1 = w1.detach().float().contiguous().to(h.device)
    w2 = w2.detach().float().contiguous().to(h.device)
    L.check(lib.diqt_se_gate(part.data_ptr(), n, nblk, vox, c, w1.shape[0], w1.data_ptr(), w2.data_ptr(), gate.data_ptr(), st), "se_gate")
    out = torch.empty_like(h)
    opart = torch.empty(n, nblk, c, 2, dtype=torch.float32, device=h.device)
    L.check(lib.diqt_scale_residual(h.data_ptr(), c, res.data_ptr(), c, out.data_ptr(), c, _dt(h), n, vox, c, gate.data_ptr(), nblk,
                                    opart.data_ptr(), 0, 0, st), "scale_residual")
    _sync()
    return (out, gate, opart, part) if return_stats else (out, gate, opart)


# ------------------------------------------------------------------ attention-block ops (csrc/attn.cu); rows = (rows, c) tensors

def chan_layernorm(x, g, beta=None, eps=1e-5, pre_act=0, res1=None, res2=None, x_sub=(0, 0)):
    """LayerNorm over the last dim of (rows, c) (imagen_pytorch3D.py:361-382); x_sub=(f, h): x rows are in merged order."""
    lib = L.load()
    rows, c = x.shape
    out = torch.empty_like(x)
    gf = g.detach().to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
    bf = beta.detach().to(device=x.device, dtype=torch.float32).contiguous() if beta is not None else None
    L.check(lib.diqt_chan_layernorm(x.data_ptr(), c, out.data_ptr(), c, _dt(x), rows, c, gf.data_ptr(), L.ptr(bf), float(eps), pre_act,
                                    L.ptr(res1), c, L.ptr(res2), c, x_sub[0], x_sub[1], L.current_stream()), "chan_layernorm")
    return out


def rows_combine(a, act=0, b=None, c2=None):
    lib = L.load()
    rows, c = a.shape
    out = torch.empty_like(a)
    L.check(lib.diqt_rows_combine(a.data_ptr(), c, act, L.ptr(b), c, L.ptr(c2), c, out.data_ptr(), c, _dt(a), rows, c, L.current_stream()), "rows_combine")
    return out


def _dw_weight(w, device):
    w = w.detach().to(device=device, dtype=torch.float32)
    return w.reshape(w.shape[0], -1).t().contiguous()


def dw_patchify(x, weight, bias, grid_dim, patch, x_sub=(0, 0)):
    """x: (rows, c) voxels of the (grid_dim*patch)^3 volume (sub-volume order if x_sub=(f, h)); weight (c, 1, p, p, p)."""
    lib = L.load()
    c = x.shape[1]
    out = torch.empty(grid_dim ** 3, c, dtype=x.dtype, device=x.device)
    w = _dw_weight(weight, x.device)
    b = bias.detach().to(device=x.device, dtype=torch.float32).contiguous() if bias is not None else None
    L.check(lib.diqt_dw_patchify(x.data_ptr(), c, out.data_ptr(), c, _dt(x), grid_dim, patch, c, w.data_ptr(), L.ptr(b), x_sub[0], x_sub[1],
                                 L.current_stream()), "dw_patchify")
    return out


def dw_conv3(x, weight, bias=None):
    """x: (d0, d1, d2, c); weight (c, 1, 3, 3, 3): depthwise 3x3x3, zero padding 1."""
    lib = L.load()
    d0, d1, d2, c = x.shape
    x = x.contiguous()
    out = torch.empty_like(x)
    w = _dw_weight(weight, x.device)
    b = bias.detach().to(device=x.device, dtype=torch.float32).contiguous() if bias is not None else None
    L.check(lib.diqt_dw_conv3(x.data_ptr(), c, out.data_ptr(), c, _dt(x), d0, d1, d2, c, w.data_ptr(), L.ptr(b), L.current_stream()), "dw_conv3")
    return out


def upsample_trilinear(tokens, grid_dim, factor):
    """tokens: (grid_dim^3, c) -> ((grid_dim*factor)^3, c), align_corners=True."""
    lib = L.load()
    c = tokens.shape[1]
    G = grid_dim * factor
    out = torch.empty(G ** 3, c, dtype=tokens.dtype, device=tokens.device)
    L.check(lib.diqt_upsample_trilinear(tokens.data_ptr(), c, out.data_ptr(), c, _dt(tokens), grid_dim, factor, c, L.current_stream()), "upsample_trilinear")
    return out


def linear_attention(qkv, heads, dim_head, act=1, impl="auto"):
    """qkv: (tokens, 3*heads*dim_head) holding q | k | v blocks -> (tokens, heads*dim_head)   (imagen_pytorch3D.py:1001-1011).
    impl "tc": the tcgen05 kernels (bf16, dim_head 64, even number of heads); "simt": the CUDA-core kernels; "auto": tc when supported."""
    lib = L.load()
    n, inner = qkv.shape[0], heads * dim_head
    esz = qkv.element_size()
    out = torch.empty(n, inner, dtype=qkv.dtype, device=qkv.device)
    p = qkv.data_ptr()
    tc_ok = bool(lib.diqt_linattn_tc_supported(_dt(qkv), dim_head, heads, 3 * inner, inner))
    if impl == "tc" and not tc_ok:
        raise L.DiqtError("linear_attention(tc): needs bf16, dim_head 64 and an even number of heads")
    if impl == "tc" or (impl == "auto" and tc_ok):
        nbytes = C.c_size_t(0)
        L.check(lib.diqt_linattn_tc_workspace_bytes(n, heads, C.byref(nbytes)), "linattn_tc_workspace_bytes")
        ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=qkv.device)
        wp = (ws.data_ptr() + 255) // 256 * 256
        plan = C.c_void_p(0)
        L.check(lib.diqt_linattn_tc_plan_create(p, p + inner * esz, p + 2 * inner * esz, 3 * inner, out.data_ptr(), inner, n, heads,
                                                float(dim_head) ** -0.5, act, wp, C.byref(plan)), "linattn_tc_plan_create")
        try:
            L.check(lib.diqt_linattn_tc_run(plan.value, L.current_stream()), "linattn_tc_run")
            _sync()
        finally:
            lib.diqt_linattn_tc_plan_destroy(plan.value)
        return out
    chunks = C.c_int(0)
    L.check(lib.diqt_linear_attention_chunks(n, C.byref(chunks)), "linear_attention_chunks")
    stat = torch.empty(inner * 2, dtype=torch.float32, device=qkv.device)
    part = torch.empty(chunks.value * heads * dim_head * dim_head, dtype=torch.float32, device=qkv.device)
    L.check(lib.diqt_linear_attention(p, p + inner * esz, p + 2 * inner * esz, 3 * inner, out.data_ptr(), inner, _dt(qkv), n, heads, dim_head,
                                      float(dim_head) ** -0.5, act, stat.data_ptr(), part.data_ptr(), L.current_stream()), "linear_attention")
    return out


def softmax_attention(qkv, heads, dim_head, act=1, impl="simt"):
    """qkv: (tokens, 3*heads*dim_head) holding q | k | v blocks -> (tokens, heads*dim_head)   (imagen_pytorch3D.py:1087-1100).
    impl "tc": the tcgen05 kernel (bf16, dim_head 64); "simt": the fp32 CUDA-core kernel."""
    lib = L.load()
    n, inner = qkv.shape[0], heads * dim_head
    esz = qkv.element_size()
    out = torch.empty(n, inner, dtype=qkv.dtype, device=qkv.device)
    p = qkv.data_ptr()
    if impl == "tc":
        if not lib.diqt_attn_tc_supported(_dt(qkv), dim_head, 3 * inner, 3 * inner, 3 * inner, inner):
            raise L.DiqtError("softmax_attention(tc): needs bf16, dim_head 64")
        nbytes = C.c_size_t(0)
        L.check(lib.diqt_attn_tc_workspace_bytes(n, heads, C.byref(nbytes)), "attn_tc_workspace_bytes")
        ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=qkv.device)
        plan = C.c_void_p(0)
        L.check(lib.diqt_attn_tc_plan_create(p, p + inner * esz, p + 2 * inner * esz, 3 * inner, 3 * inner, 3 * inner, out.data_ptr(), inner, n, heads,
                                             float(dim_head) ** -0.5, act, ws.data_ptr(), C.byref(plan)), "attn_tc_plan_create")
        try:
            L.check(lib.diqt_attn_tc_run(plan.value, L.current_stream()), "attn_tc_run")
            _sync()
        finally:
            lib.diqt_attn_tc_plan_destroy(plan.value)
        return out
    L.check(lib.diqt_softmax_attention(p, p + inner * esz, p + 2 * inner * esz, 3 * inner, 3 * inner, 3 * inner, out.data_ptr(), inner, _dt(qkv), n, heads,
                                       dim_head, float(dim_head) ** -0.5, act, L.current_stream()), "softmax_attention")
    return out
