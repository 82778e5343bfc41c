"""CPU oracle for the attention blocks of the 3-D U-Net  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Plain-PyTorch fp32 restatement of the three attention variants the reference can place after the first ResnetBlock of every
encoder level and in front of the mid block (/root/reference/imagen_pytorch3D.py:1392-1403, 1418-1430, 1610-1622, 1635-1646):

  * att_type 'linear'  : LinearAttentionTransformerBlock  :1118-1150  (LinearAttention :926-1016, ChanFeedForward :1108-1116)
  * att_type 'softmax' : SoftMaxAttentionTransformerBlock :1153-1186  (SoftMaxAttention :1018-1106)
  * att_type 'vit'     : ViT3D :871-910 (PatchEmbedding :841-856, TransformerEncoderBlock :722-746, MultiHeadAttention :811-838,
                         FeedForwardBlock :772-809)

All of them see the f^3 sub-volumes merged into one volume (utils_mine.py:44-67) and return the re-split result.
Dropout layers are identities on the sampling path (`eval_decorator`, :115-122).  Driven by a `state_dict` with the reference's
key names.  Pinned by tests/test_oracle_vs_reference.py (live reference) and tests/test_oracle_golden.py (committed fixtures).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def chan_layernorm(x: Tensor, g: Tensor) -> Tensor:
    # LayerNorm(dim=-4) (:361-382): biased variance over channels, eps 1e-5 in fp32, scale only
    var = torch.var(x, dim=-4, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=-4, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def dw_separable(sd: Dict[str, Tensor], p: str, x: Tensor, stride: int, padding: int) -> Tensor:
    # depthwise_separable_conv3d (:858-869): depthwise k^3 (groups = channels, bias) then pointwise 1x1x1 (bias)
    w = sd[p + "depthwise.weight"]
    x = F.conv3d(x, w, sd[p + "depthwise.bias"], stride=stride, padding=padding, groups=w.shape[0])
    return F.conv3d(x, sd[p + "pointwise.weight"], sd[p + "pointwise.bias"])


def _qkv(sd: Dict[str, Tensor], p: str, fmap: Tensor, heads: int):
    # to_q / to_k / to_v (:961-977): Dropout, 1x1x1 conv (no bias), depthwise 3x3x3 (no bias); then '(b h) (x y z) c'
    outs = []
    for name in ("to_q", "to_k", "to_v"):
        t = F.conv3d(fmap, sd[f"{p}{name}.1.weight"])
        w = sd[f"{p}{name}.2.weight"]
        t = F.conv3d(t, w, padding=1, groups=w.shape[0])
        b, hc, x, y, z = t.shape
        outs.append(t.reshape(b, heads, hc // heads, x * y * z).permute(0, 1, 3, 2).reshape(b * heads, x * y * z, hc // heads))
    return outs


def _attention(sd: Dict[str, Tensor], p: str, fmap: Tensor, heads: int, dim_head: int, patch_size: int, kind: str) -> Tensor:
    """LinearAttention.forward (:986-1016) / SoftMaxAttention.forward (:1078-1106) with patch=True, no context."""
    fmap = chan_layernorm(fmap, sd[p + "patch_embed.norm.g"])                               # Patchify :926-929
    fmap = dw_separable(sd, p + "patch_embed.projection.", fmap, stride=patch_size, padding=0)
    b, _, x, y, z = fmap.shape
    fmap = chan_layernorm(fmap, sd[p + "norm.g"])
    q, k, v = _qkv(sd, p, fmap, heads)
    scale = dim_head ** -0.5
    if kind == "linear":
        q = q.softmax(dim=-1) * scale
        k = k.softmax(dim=-2)
        ctx = torch.einsum("bnd,bne->bde", k, v)
        out = torch.einsum("bnd,bde->bne", q, ctx)
    else:
        energy = torch.einsum("bqd,bkd->bqk", q, k) * scale
        out = torch.einsum("bnd,bde->bne", energy.softmax(dim=-1), v)
    out = out.reshape(b, heads, x * y * z, dim_head).permute(0, 1, 3, 2).reshape(b, heads * dim_head, x, y, z)
    out = F.mish(out)
    out = chan_layernorm(F.conv3d(out, sd[p + "to_out.0.weight"]), sd[p + "to_out.1.g"])
    out = F.interpolate(out, scale_factor=patch_size, mode="trilinear", align_corners=True)  # reconstruct :952-959
    out = dw_separable(sd, p + "reconstruct.1.", out, stride=1, padding=1)
    return chan_layernorm(out, sd[p + "reconstruct.2.g"])


def chan_feedforward(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    # ChanFeedForward (:1108-1116): ChanLN, 1x1x1 (no bias), GELU (erf), ChanLN, 1x1x1 (no bias)
    h = F.conv3d(chan_layernorm(x, sd[p + "0.g"]), sd[p + "1.weight"])
    h = chan_layernorm(F.gelu(h), sd[p + "3.g"])
    return F.conv3d(h, sd[p + "4.weight"])


def attention_transformer_block(sd: Dict[str, Tensor], p: str, x: Tensor, *, depth: int, heads: int, dim_head: int,
                                patch_size: int, kind: str) -> Tensor:
    """{Linear,SoftMax}AttentionTransformerBlock.forward (:1146-1150, :1181-1186) on the merged volume (1, C, D, H, W)."""
    for i in range(depth):
        x = _attention(sd, f"{p}layers.{i}.0.", x, heads, dim_head, patch_size, kind) + x
        x = chan_feedforward(sd, f"{p}layers.{i}.1.", x) + x
    return x


def vit3d(sd: Dict[str, Tensor], p: str, x: Tensor, *, depth: int, heads: int, dim_head: int, patch_size: int,
          local: bool) -> Tensor:
    """ViT3D.forward (:905-910) on the merged volume (1, C, D, H, W)."""
    t = dw_separable(sd, p + "patch_embedding.projection.0.", x, stride=patch_size, padding=0)   # PatchEmbedding :841-856
    b, e, gh, gw, gd = t.shape
    t = t.reshape(b, e, gh * gw * gd).permute(0, 2, 1) + sd[p + "patch_embedding.positions"]
    for i in range(depth):
        q = f"{p}transformer_encoder.layers.{i}.block."
        # ResidualAdd(LayerNorm, MultiHeadAttention) :724-729, 811-838
        hdn = F.layer_norm(t, (e,), sd[q + "0.fn.0.weight"], sd[q + "0.fn.0.bias"])
        qkv = F.linear(hdn, sd[q + "0.fn.1.qkv.weight"], sd[q + "0.fn.1.qkv.bias"])
        n = qkv.shape[1]
        qkv = qkv.reshape(b, n, heads, dim_head, 3).permute(4, 0, 2, 1, 3)         # 'b n (h d qkv) -> qkv b h n d'
        energy = torch.einsum("bhqd,bhkd->bhqk", qkv[0], qkv[1]) * dim_head ** -0.5
        out = torch.einsum("bhal,bhlv->bhav", energy.softmax(dim=-1), qkv[2])
        out = out.permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
        t = F.linear(out, sd[q + "0.fn.1.projection.weight"], sd[q + "0.fn.1.projection.bias"]) + t
        # ResidualAdd(LayerNorm, FeedForwardBlock) :730-736, 772-809
        hdn = F.layer_norm(t, (e,), sd[q + "1.fn.0.weight"], sd[q + "1.fn.0.bias"])
        if local:
            v = hdn.permute(0, 2, 1).reshape(b, e, gh, gw, gd)
            # the three stages are registered twice (up_proj / depth_conv / down_proj and net.0 / net.1 / net.2, shared
            # storage); `load_state_dict` visits `net` last, so those are the keys whose values a loaded model ends up with
            v = F.mish(F.conv3d(v, sd[q + "1.fn.1.net.0.1.weight"], sd[q + "1.fn.1.net.0.1.bias"]))
            v = F.mish(dw_separable(sd, q + "1.fn.1.net.1.0.", v, stride=1, padding=1))
            v = F.conv3d(v, sd[q + "1.fn.1.net.2.0.weight"], sd[q + "1.fn.1.net.2.0.bias"])
            hdn = v.reshape(b, e, n).permute(0, 2, 1)
        else:
            hdn = F.linear(F.mish(F.linear(hdn, sd[q + "1.fn.1.net.0.weight"], sd[q + "1.fn.1.net.0.bias"])),
                           sd[q + "1.fn.1.net.3.weight"], sd[q + "1.fn.1.net.3.bias"])
        t = hdn + t
    t = F.layer_norm(t, (e,), sd[p + "reconstruction.0.weight"], sd[p + "reconstruction.0.bias"])   # :897-903
    v = t.permute(0, 2, 1).reshape(b, e, gh, gw, gd)
    v = F.interpolate(v, scale_factor=patch_size, mode="trilinear", align_corners=True)
    v = dw_separable(sd, p + "reconstruction.3.", v, stride=1, padding=1)
    return chan_layernorm(v, sd[p + "reconstruction.4.g"])
