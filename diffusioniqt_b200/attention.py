"""Parameter holders for the attention blocks of the reference U-Net
(/root/reference/imagen_pytorch3D.py:361-382, 811-1186): same attribute / index structure, hence the same
`state_dict` keys and shapes.  They hold parameters only; the arithmetic runs in libdiqt_b200.so
(`engine.UnetEngine._add_attention`, kernels in csrc/attn.cu + the conv families).
"""
from __future__ import annotations

import torch
from torch import nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - structure only
        raise RuntimeError(f"{type(self).__name__} only holds parameters; the U-Net runs through UnetEngine (CUDA kernels)")


class _Tag(_NoForward):
    """Parameter-free slot (Dropout, Mish, GELU, Rearrange, Upsample) that keeps Sequential indices aligned."""


class ChanLayerNorm(_NoForward):
    """LayerNorm(feats, dim=-4) :361-382: a single scale `g` of shape (feats, 1, 1, 1)."""

    def __init__(self, feats):
        super().__init__()
        self.g = nn.Parameter(torch.ones(feats, 1, 1, 1))


class DepthwiseSeparableConv3d(_NoForward):
    """depthwise_separable_conv3d :858-869."""

    def __init__(self, input_dim, output_dim, kernel_size, stride, padding=0):
        super().__init__()
        self.depthwise = nn.Conv3d(input_dim, input_dim, kernel_size=kernel_size, stride=stride, padding=padding, groups=input_dim)
        self.pointwise = nn.Conv3d(input_dim, output_dim, kernel_size=1)


class Patchify(_NoForward):
    """:913-924."""

    def __init__(self, in_channels, patch_size, emb_size):
        super().__init__()
        self.norm = ChanLayerNorm(in_channels)
        self.projection = DepthwiseSeparableConv3d(in_channels, emb_size, kernel_size=patch_size, stride=patch_size)


class ConvAttention(_NoForward):
    """LinearAttention :926-1016 / SoftMaxAttention :1018-1106 (identical parameters; `kind` selects the product)."""

    def __init__(self, dim, *, kind, dim_head=32, heads=8, patch_size=2):
        super().__init__()
        self.kind, self.dim, self.dim_head, self.heads, self.patch_size = kind, dim, dim_head, heads, patch_size
        self.scale = dim_head ** -0.5
        inner = dim_head * heads
        self.norm = ChanLayerNorm(dim)
        self.nonlin = _Tag()
        self.patch_embed = Patchify(dim, patch_size, dim)
        self.reconstruct = nn.Sequential(_Tag(), DepthwiseSeparableConv3d(dim, dim, kernel_size=3, stride=1, padding=1), ChanLayerNorm(dim))

        def proj():
            return nn.Sequential(_Tag(), nn.Conv3d(dim, inner, 1, bias=False), nn.Conv3d(inner, inner, 3, bias=False, padding=1, groups=inner))

        self.to_q, self.to_k, self.to_v = proj(), proj(), proj()
        self.to_context = None
        self.to_out = nn.Sequential(nn.Conv3d(inner, dim, 1, bias=False), ChanLayerNorm(dim))


def ChanFeedForward(dim, mult=2):
    """:1108-1116."""
    hidden = int(dim * mult)
    return nn.Sequential(ChanLayerNorm(dim), nn.Conv3d(dim, hidden, 1, bias=False), _Tag(), ChanLayerNorm(hidden), nn.Conv3d(hidden, dim, 1, bias=False))


class AttentionTransformerBlock(_NoForward):
    """LinearAttentionTransformerBlock :1118-1150 / SoftMaxAttentionTransformerBlock :1153-1186."""

    def __init__(self, dim, *, kind, depth=1, heads=8, dim_head=32, ff_mult=2, patch_size=2, img_size=48):
        super().__init__()
        self.kind, self.dim, self.depth, self.heads, self.dim_head, self.patch_size, self.img_size = kind, dim, depth, heads, dim_head, patch_size, img_size
        self.layers = nn.ModuleList([
            nn.ModuleList([ConvAttention(dim, kind=kind, heads=heads, dim_head=dim_head, patch_size=patch_size), ChanFeedForward(dim, ff_mult)])
            for _ in range(depth)])


# ------------------------------------------------------------------ ViT3D :871-910

class MultiHeadAttention(_NoForward):
    """:811-838."""

    def __init__(self, emb_size, num_heads, dim_head):
        super().__init__()
        self.emb_size, self.num_heads, self.dim_head = emb_size, num_heads, dim_head
        inner = dim_head * num_heads
        self.qkv = nn.Linear(emb_size, inner * 3)
        self.att_drop = _Tag()
        self.projection = nn.Linear(inner, emb_size)


class FeedForwardBlock(nn.Sequential):
    """:772-809.  The reference registers the three stages both as attributes and inside `net`, so the state_dict carries
    every tensor twice (`up_proj.1.weight` and `net.0.1.weight`, shared storage); reproduced here."""

    def __init__(self, emb_size, expansion, local):
        super().__init__()
        self.local = local
        if local:
            self.up_proj = nn.Sequential(_Tag(), nn.Conv3d(emb_size, emb_size * expansion, kernel_size=1), _Tag())
            self.depth_conv = nn.Sequential(DepthwiseSeparableConv3d(emb_size * expansion, emb_size * expansion, kernel_size=3, stride=1, padding=1), _Tag())
            self.down_proj = nn.Sequential(nn.Conv3d(emb_size * expansion, emb_size, kernel_size=1), _Tag(), _Tag())
            self.net = nn.Sequential(self.up_proj, self.depth_conv, self.down_proj)
        else:
            self.net = nn.Sequential(nn.Linear(emb_size, expansion * emb_size), _Tag(), _Tag(), nn.Linear(expansion * emb_size, emb_size))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("FeedForwardBlock only holds parameters")


class ResidualAdd(_NoForward):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class TransformerEncoderBlock(_NoForward):
    """:722-746."""

    def __init__(self, emb_size, num_heads, dim_head, forward_expansion, local):
        super().__init__()
        self.block = nn.Sequential(
            ResidualAdd(nn.Sequential(nn.LayerNorm(emb_size), MultiHeadAttention(emb_size, num_heads, dim_head), _Tag())),
            ResidualAdd(nn.Sequential(nn.LayerNorm(emb_size), FeedForwardBlock(emb_size, forward_expansion, local), _Tag())))


class TransformerEncoder(_NoForward):
    def __init__(self, depth, **kw):
        super().__init__()
        self.layers = nn.ModuleList([TransformerEncoderBlock(**kw) for _ in range(depth)])


class PatchEmbedding(_NoForward):
    """:841-856."""

    def __init__(self, in_channels, patch_size, emb_size, img_size):
        super().__init__()
        self.projection = nn.Sequential(DepthwiseSeparableConv3d(in_channels, emb_size, kernel_size=patch_size, stride=patch_size), _Tag())
        self.positions = nn.Parameter(torch.randn((img_size // patch_size) ** 3, emb_size))


class ViT3D(_NoForward):
    """:871-910."""

    def __init__(self, in_channels, patch_size, num_heads, dim_head, img_size, depth, forward_expansion, local):
        super().__init__()
        self.kind = "vit"
        self.dim, self.patch_size, self.heads, self.dim_head, self.img_size, self.depth = in_channels, patch_size, num_heads, dim_head, img_size, depth
        self.local, self.expansion = local, forward_expansion
        self.patch_embedding = PatchEmbedding(in_channels, patch_size, in_channels, img_size)
        self.transformer_encoder = TransformerEncoder(depth, emb_size=in_channels, num_heads=num_heads, dim_head=dim_head,
                                                      forward_expansion=forward_expansion, local=local)
        self.reconstruction = nn.Sequential(nn.LayerNorm(in_channels), _Tag(), _Tag(),
                                            DepthwiseSeparableConv3d(in_channels, in_channels, kernel_size=3, stride=1, padding=1),
                                            ChanLayerNorm(in_channels))


def make_attention(att_type, dim, *, patch_size, heads, dim_head, img_size, depth, ff_mult, local):
    """The block the reference builds at :1392-1403 / :1418-1430 for `att_type`."""
    if att_type == "vit":
        return ViT3D(dim, patch_size, heads, dim_head, img_size, depth, ff_mult, local)
    kind = "linear" if att_type == "linear" else "softmax"
    return AttentionTransformerBlock(dim, kind=kind, depth=depth, heads=heads, dim_head=dim_head, ff_mult=ff_mult, patch_size=patch_size, img_size=img_size)
