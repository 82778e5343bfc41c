"""CPU oracle for the 3-D U-Net forward pass  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-PyTorch (fp32, CPU, NCDHW) restatement of what the reference computes in
`Unet.forward` (/root/reference/imagen_pytorch3D.py:1554-1684) for the
configurations the reference's drivers ship (`train.py:83-116`,
`config/config.yaml`, `config/eval_config.yaml`): attention off, pixel-shuffle
upsampling, SE channel gate, optional `deep_feature` mid block, optional
`boundary` mode.  It is driven purely by a `state_dict` with the reference's key
names, so the same weights feed the reference, this
oracle and the CUDA path.

Pinning: `tests/test_oracle_vs_reference.py` compares this file with the live,
unmodified reference module (imported through `tests/golden/ref_shim.py`) when
`/root/reference` exists, and `tests/test_oracle_golden.py` compares it with the
committed fixtures in `tests/golden/*.npz` that `tests/golden/make_golden.py`
produced by running the reference itself.  The reference ships no tests and no
golden vectors of its own, so those two are the pins.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class UnetSpec:
    """The constructor arguments that change the arithmetic (imagen_pytorch3D.py:1189-1239)."""

    dim: int
    init_dim: Optional[int] = 32
    dim_mults: Sequence[int] = (1, 2, 4, 8)
    num_resnet_blocks: Sequence[int] | int = 1
    resnet_groups: Sequence[int] | int = 8
    channels: int = 3
    channels_out: Optional[int] = None
    lowres_cond: bool = False
    cond_images_channels: int = 0
    self_cond: bool = False
    learned_sinu_pos_emb_dim: int = 16
    use_se_attn: bool = True
    scale_skip_connection: bool = False
    final_resnet_block: bool = True
    deep_feature: bool = True
    boundary: bool = False
    batch_sample_factor: int = 3
    init_conv_kernel_size: int = 3
    # attention (imagen_pytorch3D.py:1392-1403, 1418-1430); off in both shipped configs
    att_type: str = "vit"
    attend_at_enc: Sequence[bool] | bool = False
    attend_at_enc_depth: Sequence[int] | int = 1
    attend_at_enc_heads: Sequence[int] | int = 8
    attend_at_middle: bool = False
    attend_at_middle_depth: int = 1
    attend_at_middle_heads: int = 8
    attn_dim_head: int = 64
    att_localvit: bool = True

    def patch_sizes(self):
        """Patch size of the attention block at each encoder level and of the mid block (:1361, 1413-1414)."""
        nl, ps, out = self.layers(), 8, []
        for l in range(nl):
            out.append(ps)
            if l != nl - 1:
                ps //= 2
        return out, ps

    def layers(self) -> int:
        return len(tuple(self.dim_mults))

    def dims(self):
        init_dim = self.init_dim if self.init_dim is not None else self.dim
        return [init_dim] + [self.dim * m for m in self.dim_mults]

    def per_layer(self, v):
        n = self.layers()
        if isinstance(v, (list, tuple)):
            assert len(v) == n
            return tuple(v)
        return (v,) * n


# ----------------------------------------------------------------------------- pieces

def mish(x: Tensor) -> Tensor:
    # nn.Mish = x * tanh(softplus(x))   (imagen_pytorch3D.py:547, 587, 1311); same ATen op as the reference
    return F.mish(x)


def merge_sub_volumes(sub: Tensor, factor: int) -> Tensor:
    """(f^3, C, a, a, a) -> (1, C, f*a, f*a, f*a); sub-volume b = i + f*j + f*f*k sits at block
    (i, j, k) of dims (2, 3, 4).  utils_mine.py:44-67 / SURVEY.md Appendix B.4."""
    B, C, a = sub.shape[0], sub.shape[1], sub.shape[2]
    f = factor
    assert B == f ** 3
    # reference: reshape(split_w, split_h, split_d, ...) then cat outermost along the LAST dim
    v = sub.reshape(f, f, f, C, a, a, a)            # [p, q, r, C, x, y, z] with b = p*f*f + q*f + r
    # p indexes blocks along dim -1, q along dim -2, r along dim -3
    v = v.permute(3, 2, 4, 1, 5, 0, 6)              # C, r, x, q, y, p, z
    return v.reshape(1, C, f * a, f * a, f * a)


def split_sub_volumes(vol: Tensor, factor: int) -> Tensor:
    """Inverse of `merge_sub_volumes` (utils_mine.py:25-42)."""
    _, C, W, _, _ = vol.shape
    f = factor
    a = W // f
    v = vol.reshape(C, f, a, f, a, f, a)            # C, r, x, q, y, p, z
    v = v.permute(5, 3, 1, 0, 2, 4, 6)              # p, q, r, C, x, y, z
    return v.reshape(f ** 3, C, a, a, a)


def boundary_pad(x: Tensor, factor: int) -> Tensor:
    """imagen_pytorch3D.py:37-46: merge the f^3 sub-volumes, zero-pad the merged volume by one
    voxel, and cut it again into sub-volumes that carry a one-voxel halo of their neighbours."""
    B, C, h = x.shape[0], x.shape[1], x.shape[2]
    f = factor
    big = F.pad(merge_sub_volumes(x, f), (1, 1, 1, 1, 1, 1))
    out = x.new_empty(B, C, h + 2, h + 2, h + 2)
    for p in range(f):
        for q in range(f):
            for r in range(f):
                b = p * f * f + q * f + r
                out[b] = big[0, :, r * h:r * h + h + 2, q * h:q * h + h + 2, p * h:p * h + h + 2]
    return out


def conv3(x: Tensor, w: Tensor, b: Tensor, spec: UnetSpec) -> Tensor:
    # Block.project (imagen_pytorch3D.py:550-553, 563-566)
    if spec.boundary:
        return F.conv3d(boundary_pad(x, spec.batch_sample_factor), w, b)
    return F.conv3d(x, w, b, padding=1)


def block(sd: Dict[str, Tensor], p: str, x: Tensor, groups: int, spec: UnetSpec,
          scale_shift: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    # Block.forward (imagen_pytorch3D.py:555-566)
    x = F.group_norm(x, groups, sd[p + "groupnorm.weight"], sd[p + "groupnorm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    x = mish(x)
    return conv3(x, sd[p + "project.weight"], sd[p + "project.bias"], spec)


def se_gate(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    # SE3D.forward (imagen_pytorch3D.py:617-632); both Linear layers are bias-free
    y = x.mean(dim=(2, 3, 4))
    y = torch.relu(F.linear(y, sd[p + "fc.0.weight"]))
    y = torch.sigmoid(F.linear(y, sd[p + "fc.2.weight"]))
    return x * y[:, :, None, None, None]


def resnet_block(sd: Dict[str, Tensor], p: str, x: Tensor, t: Tensor, groups: int, spec: UnetSpec) -> Tensor:
    # ResnetBlock.forward (imagen_pytorch3D.py:600-614).  FiLM goes to block2 only (:607-608).
    te = F.linear(mish(t), sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
    scale, shift = te[:, :, None, None, None].chunk(2, dim=1)
    h = block(sd, p + "block1.", x, groups, spec)
    h = block(sd, p + "block2.", h, groups, spec, (scale, shift))
    if (p + "se.fc.0.weight") in sd:
        h = se_gate(sd, p + "se.", h)
    if (p + "res_conv.weight") in sd:
        x = F.conv3d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def pixel_unshuffle3d(x: Tensor) -> Tensor:
    # Rearrange('b c (h s1) (w s2) (d s3) -> b (c s1 s2 s3) h w d')  (imagen_pytorch3D.py:489-496)
    B, C, H, W, D = x.shape
    x = x.reshape(B, C, H // 2, 2, W // 2, 2, D // 2, 2)
    x = x.permute(0, 1, 3, 5, 7, 2, 4, 6)
    return x.reshape(B, C * 8, H // 2, W // 2, D // 2)


def pixel_shuffle3d(x: Tensor) -> Tensor:
    # PixelShuffle3D(2)  (imagen_pytorch3D.py:416-439): out[c, 2d+i, 2h+j, 2w+k] = in[c*8+4i+2j+k, d, h, w]
    B, C8, D, H, W = x.shape
    C = C8 // 8
    x = x.reshape(B, C, 2, 2, 2, D, H, W)
    x = x.permute(0, 1, 5, 2, 6, 3, 7, 4)
    return x.reshape(B, C, D * 2, H * 2, W * 2)


def attention_block(sd: Dict[str, Tensor], p: str, x: Tensor, spec: UnetSpec, *, depth: int, heads: int, patch_size: int) -> Tensor:
    """merge the f^3 sub-volumes, run the attention block on the merged volume, split again (:1613-1617, 1638-1642)."""
    from . import attn_oracle as A
    f = spec.batch_sample_factor
    v = merge_sub_volumes(x, f)
    if spec.att_type == "vit":
        v = A.vit3d(sd, p, v, depth=depth, heads=heads, dim_head=spec.attn_dim_head, patch_size=patch_size, local=spec.att_localvit)
    else:
        v = A.attention_transformer_block(sd, p, v, depth=depth, heads=heads, dim_head=spec.attn_dim_head, patch_size=patch_size,
                                          kind="linear" if spec.att_type == "linear" else "softmax")
    return split_sub_volumes(v, f)


def time_embedding(sd: Dict[str, Tensor], time: Tensor) -> Tensor:
    """to_time_hiddens + to_time_cond (imagen_pytorch3D.py:518-533, 1305-1316, 1597-1599)."""
    w = sd["to_time_hiddens.0.weights"]
    tcol = time[:, None].to(torch.float32)
    freqs = tcol * w[None, :] * 2 * math.pi
    four = torch.cat((tcol, freqs.sin(), freqs.cos()), dim=-1)
    hid = mish(F.linear(four, sd["to_time_hiddens.1.weight"], sd["to_time_hiddens.1.bias"]))
    return F.linear(hid, sd["to_time_cond.0.weight"], sd["to_time_cond.0.bias"])


# ----------------------------------------------------------------------------- forward

def unet_forward(sd: Dict[str, Tensor], spec: UnetSpec, x: Tensor, time: Tensor, *,
                 lowres_cond_img: Optional[Tensor] = None, cond_images: Optional[Tensor] = None,
                 self_cond: Optional[Tensor] = None, taps: Optional[dict] = None) -> Tensor:
    """x, lowres_cond_img: (B, C, S, S, S) fp32; time: (B,) log-SNR.  Returns (B, C_out, S, S, S).

    `taps`, if given, is filled with named intermediate activations for per-kernel parity tests.
    """
    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    nl = spec.layers()
    nblocks = spec.per_layer(spec.num_resnet_blocks)
    groups = spec.per_layer(spec.resnet_groups)

    if spec.self_cond:                                                       # :1569-1571
        x = torch.cat((x, self_cond if self_cond is not None else torch.zeros_like(x)), dim=1)
    assert not (spec.lowres_cond and lowres_cond_img is None)                # :1573
    if lowres_cond_img is not None:
        x = torch.cat((x, lowres_cond_img), dim=1)                           # :1576
    assert (spec.cond_images_channels > 0) == (cond_images is not None)      # :1579
    if cond_images is not None:
        x = torch.cat((cond_images, x), dim=1)                               # :1584

    k = spec.init_conv_kernel_size
    if "init_conv.convs.0.weight" in sd:                                     # CrossEmbedLayer (:661-686, init_cross_embed=True)
        assert not spec.boundary
        maps, i = [], 0
        while f"init_conv.convs.{i}.weight" in sd:
            w = sd[f"init_conv.convs.{i}.weight"]
            maps.append(F.conv3d(x, w, sd[f"init_conv.convs.{i}.bias"], padding=(w.shape[-1] - 1) // 2))
            i += 1
        x = torch.cat(maps, dim=1)
    elif spec.boundary:                                                        # :1587-1589
        x = F.conv3d(boundary_pad(x, spec.batch_sample_factor), sd["init_conv.weight"], sd["init_conv.bias"])
    else:
        x = F.conv3d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=k // 2)
    tap("init_conv", x)

    t = time_embedding(sd, time)
    tap("time_cond", t)

    # the reference indexes attend_at_enc[ind] (:1392): longer sequences than the level count are legal
    att_enc = tuple(spec.attend_at_enc)[:nl] if isinstance(spec.attend_at_enc, (list, tuple)) else (spec.attend_at_enc,) * nl
    att_depth = spec.per_layer(spec.attend_at_enc_depth)
    att_heads = spec.per_layer(spec.attend_at_enc_heads)
    enc_ps, mid_ps = spec.patch_sizes()

    hiddens = []
    for l in range(nl):                                                      # :1604-1631
        x = resnet_block(sd, f"downs.{l}.1.", x, t, groups[l], spec)
        tap(f"downs.{l}.1", x)
        if att_enc[l]:                                                       # :1610-1622 (x += res)
            a = attention_block(sd, f"downs.{l}.2.", x, spec, depth=att_depth[l], heads=att_heads[l], patch_size=enc_ps[l])
            tap(f"downs.{l}.2", merge_sub_volumes(a, spec.batch_sample_factor))   # what a forward hook on the module sees
            x = a + x
        for i in range(nblocks[l]):
            x = resnet_block(sd, f"downs.{l}.3.{i}.", x, t, groups[l], spec)
            tap(f"downs.{l}.3.{i}", x)
        if l != nl - 1:
            hiddens.append(x)
            x = F.conv3d(pixel_unshuffle3d(x), sd[f"downs.{l}.4.1.weight"], sd[f"downs.{l}.4.1.bias"])
        else:
            x = F.conv3d(x, sd[f"downs.{l}.4.weight"], sd[f"downs.{l}.4.bias"])  # :1388
        tap(f"downs.{l}.4", x)

    if spec.deep_feature:                                                    # :1633-1651
        if spec.attend_at_middle:                                            # :1635-1646 (no residual around mid_attn)
            x = attention_block(sd, "mid_attn.", x, spec, depth=spec.attend_at_middle_depth, heads=spec.attend_at_middle_heads,
                                patch_size=mid_ps)
            tap("mid_attn", merge_sub_volumes(x, spec.batch_sample_factor))
        x = resnet_block(sd, "mid_block.", x, t, groups[-1], spec)
        tap("mid_block", x)

    skip_scale = 1.0 if not spec.scale_skip_connection else 2 ** -0.5        # :1346
    rgroups = tuple(reversed(groups))
    rblocks = tuple(reversed(nblocks))
    for u in range(nl):                                                      # :1657-1664
        last = u == nl - 1
        if not last:
            if f"ups.{u}.0.deconv.0.weight" in sd:                           # Upsample_deconv (:440-457), pixel_shuffle_upsample=False
                x = mish(F.conv_transpose3d(x, sd[f"ups.{u}.0.deconv.0.weight"], sd[f"ups.{u}.0.deconv.0.bias"], stride=2, padding=1, output_padding=1))
            else:
                w, b = sd[f"ups.{u}.0.net.0.weight"], sd[f"ups.{u}.0.net.0.bias"]
                x = pixel_shuffle3d(mish(F.conv3d(x, w, b)))                 # :459-487
            tap(f"ups.{u}.0", x)
            x = torch.cat((x, hiddens.pop() * skip_scale), dim=1)            # :1653
        x = resnet_block(sd, f"ups.{u}.1.", x, t, rgroups[u], spec)
        tap(f"ups.{u}.1", x)
        for i in range(rblocks[u]):
            x = resnet_block(sd, f"ups.{u}.2.{i}.", x, t, rgroups[u], spec)
            tap(f"ups.{u}.2.{i}", x)

    if spec.final_resnet_block:                                              # :1677
        x = resnet_block(sd, "final_res_block.", x, t, groups[0], spec)
        tap("final_res_block", x)
    out = F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"])        # :1682
    return tap("final_conv", out)


def count_flops(spec: UnetSpec, batch: int, size: int) -> float:
    """Algorithmic FLOPs of one forward, counted the way BASELINE.md section 3 does: Conv3d
    2*Cin*Cout*k^3*voxels and Linear 2*in*out*rows, nothing else."""
    dims = spec.dims()
    nl = spec.layers()
    nblocks = spec.per_layer(spec.num_resnet_blocks)
    tdim = spec.dim * 4
    fl = 0.0

    def conv(ci, co, k, s):
        return 2.0 * ci * co * k ** 3 * batch * s ** 3

    def lin(i, o):
        return 2.0 * i * o * batch

    def res(ci, co, s, se=True):
        f = lin(tdim, 2 * co) + conv(ci, co, 3, s) + conv(co, co, 3, s)
        if se:
            f += lin(co, co // 16) + lin(co // 16, co)
        if ci != co:
            f += conv(ci, co, 1, s)
        return f

    cin = spec.channels * (1 + int(spec.lowres_cond)) + spec.cond_images_channels + (spec.channels if spec.self_cond else 0)
    fl += conv(cin, dims[0], spec.init_conv_kernel_size, size)
    fl += lin(spec.learned_sinu_pos_emb_dim + 1, tdim) + lin(tdim, tdim)
    s = size
    skips = []
    for l in range(nl):
        di, do = dims[l], dims[l + 1]
        fl += (1 + nblocks[l]) * res(di, di, s, spec.use_se_attn)
        if l != nl - 1:
            skips.append(di)
            s //= 2
            fl += conv(di * 8, do, 1, s)
        else:
            fl += conv(di, do, 1, s)
    if spec.deep_feature:
        fl += res(dims[-1], dims[-1], s, False)
    rb = tuple(reversed(nblocks))
    x_dim = dims[-1]
    for u in range(nl):
        do, _ = list(reversed(list(zip(dims[:-1], dims[1:]))))[u]
        last = u == nl - 1
        if not last:
            fl += conv(x_dim, do * 8, 1, s)
            s *= 2
            fl += res(do + skips.pop(), do, s, spec.use_se_attn)
        else:
            fl += res(x_dim, do, s, spec.use_se_attn)
        fl += rb[u] * res(do, do, s, spec.use_se_attn)
        x_dim = do
    if spec.final_resnet_block:
        fl += res(x_dim, spec.dim, s, spec.use_se_attn)
        x_dim = spec.dim
    fl += conv(x_dim, spec.channels_out or spec.channels, 1, s)
    return fl
