"""Host-side mirror of the reference 3-D U-Net (`Unet`, `SRUnet256`, ... in
/root/reference/imagen_pytorch3D.py:1188-1738).

Same constructor keywords, `forward` / `forward_with_cond_scale` signatures, persistence helpers
and - most importantly - the same `state_dict` key names and shapes, so
a reference checkpoint loads unchanged.  The modules below only *hold* parameters; all arithmetic
runs in libdiqt_b200.so through `diffusioniqt_b200.engine.UnetEngine` (hand-written sm_100a
kernels, no PyTorch fallback).
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Optional

import torch
from torch import nn

from .attention import make_attention


def _cast_tuple(val, length=None):
    if isinstance(val, list):
        val = tuple(val)
    out = val if isinstance(val, tuple) else ((val,) * (length if length is not None else 1))
    if length is not None:
        assert len(out) == length
    return out


# ------------------------------------------------------------------ parameter holders

class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - structure only
        raise RuntimeError(f"{type(self).__name__} only holds parameters; the U-Net runs through UnetEngine (CUDA kernels)")


class LearnedSinusoidalPosEmb(_NoForward):
    """imagen_pytorch3D.py:518-533 (parameter name `weights`)."""

    def __init__(self, dim):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))


class _Tag(_NoForward):
    """Parameter-free slot (nn.Mish, Rearrange, PixelShuffle3D, ...) that keeps Sequential indices aligned."""


class Block(_NoForward):
    """GroupNorm -> [FiLM] -> Mish -> Conv3d 3x3x3 (imagen_pytorch3D.py:535-566)."""

    def __init__(self, dim, dim_out, groups=8, boundary=False):
        super().__init__()
        self.groupnorm = nn.GroupNorm(groups, dim)
        self.project = nn.Conv3d(dim, dim_out, 3) if boundary else nn.Conv3d(dim, dim_out, 3, padding=1)


class SE3D(_NoForward):
    """imagen_pytorch3D.py:617-632: fc = [Linear(c, c/r, no bias), ReLU, Linear(c/r, c, no bias), Sigmoid]."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), _Tag(),
                                nn.Linear(channel // reduction, channel, bias=False), _Tag())


class ResnetBlock(_NoForward):
    """imagen_pytorch3D.py:568-614."""

    def __init__(self, dim, dim_out, time_cond_dim=None, groups=8, use_se=False, boundary=False):
        super().__init__()
        self.dim, self.dim_out, self.groups = dim, dim_out, groups
        self.time_mlp = nn.Sequential(_Tag(), nn.Linear(time_cond_dim, dim_out * 2)) if time_cond_dim is not None else None
        self.block1 = Block(dim, dim_out, groups=groups, boundary=boundary)
        self.block2 = Block(dim_out, dim_out, groups=groups, boundary=boundary)
        self.se = SE3D(dim_out, reduction=16) if use_se else _Tag()
        self.res_conv = nn.Conv3d(dim, dim_out, 1) if dim != dim_out else _Tag()

    @property
    def has_se(self):
        return isinstance(self.se, SE3D)

    @property
    def has_res_conv(self):
        return isinstance(self.res_conv, nn.Conv3d)


class PixelShuffleUpsample(_NoForward):
    """imagen_pytorch3D.py:459-487: net = [Conv3d(dim, 8*dim_out, 1), Mish, PixelShuffle3D(2)]."""

    def __init__(self, dim, dim_out=None):
        super().__init__()
        dim_out = dim_out if dim_out is not None else dim
        conv = nn.Conv3d(dim, dim_out * 8, 1)
        self.net = nn.Sequential(conv, _Tag(), _Tag())
        # ICNR-style init: groups of 4 consecutive output channels share one kernel (:477-484)
        o, i = conv.weight.shape[:2]
        w = torch.empty(o // 4, i, 1, 1, 1)
        nn.init.kaiming_uniform_(w)
        conv.weight.data.copy_(w.repeat_interleave(4, dim=0))
        nn.init.zeros_(conv.bias.data)


class Deconv3D(_NoForward):
    """imagen_pytorch3D.py:440-457 (`Upsample_deconv`): deconv = [ConvTranspose3d(k 3, stride 2, padding 1, output_padding 1), Mish]."""

    def __init__(self, inp_feat, out_feat):
        super().__init__()
        self.deconv = nn.Sequential(nn.ConvTranspose3d(inp_feat, out_feat, kernel_size=3, stride=2, padding=1, output_padding=1, bias=True), _Tag())


class CrossEmbedLayer(_NoForward):
    """imagen_pytorch3D.py:661-686: parallel convs of several kernel sizes over the same input, concatenated along channels."""

    def __init__(self, dim_in, kernel_sizes, dim_out=None, stride=2):
        super().__init__()
        assert all((t % 2) == (stride % 2) for t in kernel_sizes)
        dim_out = dim_out if dim_out is not None else dim_in
        kernel_sizes = sorted(kernel_sizes)
        num_scales = len(kernel_sizes)
        dim_scales = [int(dim_out / (2 ** i)) for i in range(1, num_scales)]
        dim_scales = [*dim_scales, dim_out - sum(dim_scales)]
        self.kernel_sizes, self.dim_scales, self.stride = kernel_sizes, dim_scales, stride
        self.convs = nn.ModuleList([nn.Conv3d(dim_in, ds, k, stride=stride, padding=(k - stride) // 2) for k, ds in zip(kernel_sizes, dim_scales)])


def Downsample(dim, dim_out=None):
    """pixel-unshuffle (parameter-free, index 0) + 1x1x1 conv (index 1)  (imagen_pytorch3D.py:489-496)."""
    dim_out = dim_out if dim_out is not None else dim
    return nn.Sequential(_Tag(), nn.Conv3d(dim * 8, dim_out, 1))


# ------------------------------------------------------------------ the U-Net

class Unet(nn.Module):
    """Drop-in for `imagen_pytorch3D.Unet` (sampling path).  See module docstring."""

    def __init__(
        self,
        *,
        dim,
        img_size=96,
        num_resnet_blocks=1,
        cond_dim=None,
        learned_sinu_pos_emb_dim=16,
        dim_mults=(1, 2, 4, 8),
        cond_images_channels=0,
        channels=3,
        channels_out=None,
        attn_dim_head=64,
        attn_heads=8,
        ff_mult=2.,
        lowres_cond=False,
        att_type='vit',
        attend_at_middle=True,
        attend_at_middle_depth=1,
        attend_at_middle_heads=8,
        attend_at_enc=True,
        attend_at_enc_depth=1,
        attend_at_enc_heads=8,
        att_drop=0.1,
        att_forward_drop=0.3,
        att_forward_expansion=2,
        att_skip_scale=False,
        att_localvit=True,
        groups=1,
        emb_size=768,
        init_dim=32,
        resnet_groups=8,
        init_conv_kernel_size=3,
        init_cross_embed=True,
        init_cross_embed_kernel_sizes=(3, 7, 15),
        cross_embed_downsample=False,
        cross_embed_downsample_kernel_sizes=(2, 4),
        memory_efficient=False,
        init_conv_to_final_conv_residual=False,
        use_se_attn=True,
        scale_skip_connection=False,
        final_resnet_block=True,
        final_conv_kernel_size=1,
        self_cond=False,
        combine_upsample_fmaps=False,
        pixel_shuffle_upsample=True,
        boundary=False,
        batch_sample=True,
        batch_sample_factor=3,
        deep_feature=True,
    ):
        super().__init__()
        self._locals = dict(locals())
        self._locals.pop('self', None)
        self._locals.pop('__class__', None)

        assert attn_heads > 1, 'you need to have more than 1 attention head, ideally at least 4 or 8'

        num_layers = len(tuple(dim_mults))
        # the reference indexes attend_at_enc[ind] per level (:1398): a bool breaks there, longer sequences are fine
        attend_at_enc = tuple(attend_at_enc)[:num_layers] if isinstance(attend_at_enc, (list, tuple)) else (attend_at_enc,) * num_layers
        # -------- options of the reference this build does not take
        unsupported = []
        if init_cross_embed and boundary:
            unsupported.append("init_cross_embed=True with boundary=True (boundary_pad followed by padded convs changes the volume size in the reference, :1587-1589)")
        if cross_embed_downsample:
            unsupported.append("cross_embed_downsample=True (cannot be constructed in the reference either: `downsample_klass(current_dim, dim_out)` "
                               "passes dim_out into CrossEmbedLayer's `kernel_sizes` slot that the partial at :1342 already fills -> TypeError at :1388)")
        if memory_efficient:
            unsupported.append("memory_efficient=True (broken in the reference: changes the output size; note that SRUnet256 / SRUnet1024 keep the "
                               "reference's default memory_efficient=True, so pass memory_efficient=False explicitly, as train.py:100 and test_all.py do)")
        if not pixel_shuffle_upsample and boundary:
            unsupported.append("pixel_shuffle_upsample=False (ConvTranspose3d upsampling) with boundary=True")
        if init_conv_to_final_conv_residual:
            unsupported.append("init_conv_to_final_conv_residual=True (channel mismatch in the reference itself)")
        has_attn = any(attend_at_enc) or (deep_feature and attend_at_middle)
        if has_attn and attn_dim_head not in (16, 32, 64):
            unsupported.append(f"attention with attn_dim_head={attn_dim_head} (kernels take 16, 32 or 64)")
        if self_cond:
            unsupported.append("self_cond=True (the reference never widens init_conv for it, imagen_pytorch3D.py:1273-1286, so it cannot run there either)")
        if init_conv_kernel_size != 3:
            unsupported.append("init_conv_kernel_size != 3")
        if final_conv_kernel_size != 1:
            unsupported.append("final_conv_kernel_size != 1")
        if unsupported:
            raise NotImplementedError("diffusioniqt_b200.Unet does not implement: " + "; ".join(unsupported))

        self.att_type = att_type
        self.dim_head = attn_dim_head
        self.att_localvit = att_localvit
        self.att_forward_expansion = att_forward_expansion
        self.batch_sample = batch_sample
        self.img_size = img_size
        self.batch_sample_factor = batch_sample_factor
        self.boundary = boundary
        self.deep_feature = deep_feature
        self.channels = channels
        self.channels_out = channels_out if channels_out is not None else channels
        self.self_cond = self_cond
        self.lowres_cond = lowres_cond
        self.has_cond_image = cond_images_channels > 0
        self.cond_images_channels = cond_images_channels
        self.learned_sinu_pos_emb_dim = learned_sinu_pos_emb_dim

        init_channels = channels * (1 + int(lowres_cond)) + cond_images_channels
        # NB: the reference does not widen init_conv for self_cond (imagen_pytorch3D.py:1273-1286)
        init_dim = init_dim if init_dim is not None else dim
        self.init_channels = init_channels
        if init_cross_embed:    # the constructor default (:1222-1223, 1289-1291); the shipped drivers pass False (train.py:91)
            self.init_conv = CrossEmbedLayer(init_channels, init_cross_embed_kernel_sizes, dim_out=init_dim, stride=1)
        else:
            self.init_conv = nn.Conv3d(init_channels, init_dim, 3) if boundary else nn.Conv3d(init_channels, init_dim, 3, padding=1)

        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.dims, self.in_out = dims, in_out

        cond_dim = cond_dim if cond_dim is not None else dim
        time_cond_dim = dim * 4
        self.time_cond_dim = time_cond_dim
        self.to_time_hiddens = nn.Sequential(LearnedSinusoidalPosEmb(learned_sinu_pos_emb_dim),
                                             nn.Linear(learned_sinu_pos_emb_dim + 1, time_cond_dim), _Tag())
        self.to_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
        self.norm_cond = nn.LayerNorm(cond_dim)  # present in checkpoints, unused on this path
        self.text_to_cond = None

        num_resnet_blocks = _cast_tuple(num_resnet_blocks, num_layers)
        resnet_groups = _cast_tuple(resnet_groups, num_layers)
        self.num_resnet_blocks, self.resnet_groups = num_resnet_blocks, resnet_groups
        self.skip_connect_scale = 1. if not scale_skip_connection else (2 ** -0.5)

        rb = dict(time_cond_dim=time_cond_dim, boundary=boundary)
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        skip_dims = []
        # attention geometry (:1361, 1376-1379, 1412-1414): patch size 8 at the first level, halved per level (not after the last);
        # img_size halves per level and only sizes the ViT position table
        att_depth = _cast_tuple(attend_at_enc_depth, num_layers)
        att_heads = _cast_tuple(attend_at_enc_heads, num_layers)
        att_kw = dict(dim_head=attn_dim_head, ff_mult=att_forward_expansion, local=att_localvit)
        patch_size, level_img = 8, img_size
        for ind, ((dim_in, dim_out), nblk, grp) in enumerate(zip(in_out, num_resnet_blocks, resnet_groups)):
            is_last = ind >= num_layers - 1
            if not is_last:
                skip_dims.append(dim_in)
            post = Downsample(dim_in, dim_out) if not is_last else nn.Conv3d(dim_in, dim_out, 1)
            attn = make_attention(att_type, dim_in, patch_size=patch_size, heads=att_heads[ind], img_size=level_img, depth=att_depth[ind],
                                  **att_kw) if attend_at_enc[ind] else None
            last_img = level_img
            level_img //= 2
            if not is_last:
                patch_size //= 2
            self.downs.append(nn.ModuleList([
                None,
                ResnetBlock(dim_in, dim_in, groups=grp, use_se=use_se_attn, **rb),
                attn,
                nn.ModuleList([ResnetBlock(dim_in, dim_in, groups=grp, use_se=use_se_attn, **rb) for _ in range(nblk)]),
                post,
            ]))

        mid_dim = dims[-1]
        if deep_feature:
            self.mid_attn = make_attention(att_type, mid_dim, patch_size=patch_size, heads=attend_at_middle_heads, img_size=last_img,
                                           depth=attend_at_middle_depth, **att_kw) if attend_at_middle else None
        self.mid_block = ResnetBlock(mid_dim, mid_dim, groups=resnet_groups[-1], **rb)  # no SE (:1432-1434)

        for ind, ((dim_out, dim_in), nblk, grp) in enumerate(zip(reversed(in_out), reversed(num_resnet_blocks), reversed(resnet_groups))):
            if ind == 0:
                dim_in = mid_dim
            is_last = ind == num_layers - 1
            if not is_last:
                skip = skip_dims.pop()
            self.ups.append(nn.ModuleList([
                (PixelShuffleUpsample(dim_in, dim_out) if pixel_shuffle_upsample else Deconv3D(dim_in, dim_out)) if not is_last else None,
                ResnetBlock(dim_out + skip, dim_out, groups=grp, use_se=use_se_attn, **rb) if not is_last
                else ResnetBlock(dim_in, dim_out, groups=grp, use_se=use_se_attn, **rb),
                nn.ModuleList([ResnetBlock(dim_out, dim_out, groups=grp, use_se=use_se_attn, **rb) for _ in range(nblk)]),
            ]))

        self.init_conv_to_final_conv_residual = init_conv_to_final_conv_residual
        final_conv_dim = dim_out
        self.final_res_block = ResnetBlock(final_conv_dim, dim, groups=resnet_groups[0], use_se=use_se_attn, **rb) if final_resnet_block else None
        self.final_conv = nn.Conv3d(dim if final_resnet_block else final_conv_dim, self.channels_out, final_conv_kernel_size)

        # ---- execution state (not part of the checkpoint contract)
        self.compute_dtype = 'bf16'   # 'bf16' (tcgen05 tensor cores) or 'fp32' (exact mode)
        self.conv_impl = 'auto'       # 'auto' | 'simt' | 'tc'
        self._engines = {}

    # ------------------------------------------------------------------ reference helper API
    def cast_model_parameters(self, *, lowres_cond, channels, channels_out):
        # imagen_pytorch3D.py:1482-1500
        if lowres_cond == self.lowres_cond and channels == self.channels and channels_out == self.channels_out:
            return self
        updated = dict(lowres_cond=lowres_cond, channels=channels, channels_out=channels_out)
        new = self.__class__(**{**self._locals, **updated})
        new.compute_dtype, new.conv_impl = self.compute_dtype, self.conv_impl
        return new

    def to_config_and_state_dict(self):
        return self._locals, self.state_dict()

    @classmethod
    def from_config_and_state_dict(klass, config, state_dict):
        unet = klass(**config)
        unet.load_state_dict(state_dict)
        return unet

    def persist_to_file(self, path):
        path = Path(path)
        path.parents[0].mkdir(exist_ok=True, parents=True)
        config, state_dict = self.to_config_and_state_dict()
        torch.save(dict(config=config, state_dict=state_dict), str(path))

    @classmethod
    def hydrate_from_file(klass, path):
        path = Path(path)
        assert path.exists()
        pkg = torch.load(str(path), weights_only=False)
        assert 'config' in pkg and 'state_dict' in pkg
        return Unet.from_config_and_state_dict(pkg['config'], pkg['state_dict'])

    # ------------------------------------------------------------------ engine management
    def set_compute_dtype(self, dtype: str):
        assert dtype in ('bf16', 'fp32')
        if dtype != self.compute_dtype:
            self.compute_dtype = dtype
            self.invalidate_engines()
        return self

    def invalidate_engines(self):
        for e in self._engines.values():
            e.close()
        self._engines = {}

    refresh_weights = invalidate_engines   # explicit name for "I changed parameters behind autograd's back (p.data[...] = ...)"

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_engines()  # packed weight copies would be stale
        return super().load_state_dict(*args, **kwargs)

    def _param_signature(self):
        return tuple((p.data_ptr(), p.device, p.dtype, p._version) for p in self.parameters())

    def _apply(self, fn, *args, **kwargs):
        # `.to()` / `.cuda()` / `.float()` land here.  Engines hold packed copies of the weights, so they are dropped when a parameter
        # really moved or changed -- and ONLY then: Imagen.sample calls `.to(device)` on every call (imagen_pytorch3D.py:1941-1962),
        # which must not throw away the compiled engine and its captured CUDA graphs.
        before = self._param_signature()
        out = super()._apply(fn, *args, **kwargs)
        if self._param_signature() != before:
            self.invalidate_engines()
        return out

    def engine_for(self, batch, dims, device):
        from .engine import UnetEngine
        taps = tuple(getattr(self, "debug_taps", ()) or ())
        key = (int(batch), tuple(int(d) for d in dims), self.compute_dtype, self.conv_impl, str(device), taps)
        # engines hold packed copies of the weights: an in-place parameter update (optimizer / EMA step, p.copy_()) bumps `_version`,
        # a re-assigned `.data` changes `data_ptr`; either drops every engine (and its captured graphs).  Writes through `.data` that
        # keep the storage are invisible to both: call invalidate_engines() / refresh_weights() after those.
        sig = self._param_signature()
        if self._engines and sig != getattr(self, "_engine_sig", None):
            self.invalidate_engines()
        self._engine_sig = sig
        eng = self._engines.get(key)
        if eng is None:
            eng = UnetEngine(self, key[0], key[1], dtype=self.compute_dtype, device=device, conv_impl=self.conv_impl, taps=taps)
            self._engines[key] = eng
        return eng

    # ------------------------------------------------------------------ forward
    def forward_with_cond_scale(self, *args, cond_scale=1., **kwargs):
        # imagen_pytorch3D.py:1540-1552
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        null_logits = self.forward(*args, cond_drop_prob=1., **kwargs)
        return null_logits + (logits - null_logits) * cond_scale

    @torch.no_grad()
    def forward(self, x, time_steps, time, *, lowres_cond_img=None, cond_images=None, self_cond=None, cond_drop_prob=0.):
        """x: (B, C, S0, S1, S2); `time_steps` is ignored (as in the reference); `time`: (B,) log-SNR.
        Inference only: returns a new fp32 tensor (B, C_out, S0, S1, S2) and records no autograd graph."""
        assert not (self.lowres_cond and lowres_cond_img is None), 'low resolution conditioning image must be present'
        assert not (self.has_cond_image ^ (cond_images is not None)), \
            'you either requested to condition on an image on the unet, but the conditioning image is not supplied, or vice versa'
        if cond_images is not None:
            assert cond_images.shape[1] == self.cond_images_channels, \
                'the number of channels on the conditioning image you are passing in does not match what you specified on initialiation of the unet'
        if not x.is_cuda:
            raise RuntimeError("diffusioniqt_b200.Unet runs only on a CUDA device (sm_100a kernels; there is no CPU fallback)")
        eng = self.engine_for(x.shape[0], x.shape[2:], x.device)
        return eng.forward(x, time, lowres_cond_img=lowres_cond_img, cond_images=cond_images, self_cond=self_cond)


Unet3D = Unet  # BASELINE.json's north star calls the 3-D U-Net "Unet3D"


class NullUnet(nn.Module):
    """imagen_pytorch3D.py:1688-1699."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.lowres_cond = False
        self.dummy_parameter = nn.Parameter(torch.tensor([0.]))

    def cast_model_parameters(self, *args, **kwargs):
        return self

    def forward(self, x, *args, **kwargs):
        return x


class BaseUnet64(Unet):
    def __init__(self, *args, **kwargs):
        default_kwargs = dict(dim=512, dim_mults=(1, 2, 3, 4), num_resnet_blocks=3, attn_heads=8, ff_mult=2., memory_efficient=False)
        super().__init__(*args, **{**default_kwargs, **kwargs})


class SRUnet256(Unet):
    def __init__(self, *args, **kwargs):
        default_kwargs = dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 8, 8), attn_heads=8, ff_mult=2., memory_efficient=True)
        super().__init__(*args, **{**default_kwargs, **kwargs})


class SRUnet1024(Unet):
    def __init__(self, *args, **kwargs):
        default_kwargs = dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 8, 8), attn_heads=8, ff_mult=2., memory_efficient=True)
        kwargs.pop('layer_cross_attns', None)
        super().__init__(*args, **{**default_kwargs, **kwargs})
