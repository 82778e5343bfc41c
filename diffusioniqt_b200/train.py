"""Training step of the 3-D U-Net on the sm_100a kernels: the forward pass with its activations kept, the loss, and the hand-written
reverse pass that `loss.backward()` runs in the reference through PyTorch autograd.

Reference call sites: `Imagen.forward` / `p_losses` (imagen_pytorch3D.py:2277-2387, 2389-2460), the autograd graph of `Unet.forward`
(:1554-1684: `ResnetBlock` :568-614, `Block` :535-566, `SE3D` :617-632, `Downsample` :489-496, `PixelShuffleUpsample` :459-487) and
`ImagenTrainer.forward / update` (trainer.py:1038-1130).

Scope: the configuration the shipped drivers train (train.py:83-116): attention off, pixel-shuffle upsampling, SE gate, plain init conv,
`boundary` off, `deep_feature` on or off.  Other branches raise `NotImplementedError`.

Where the arithmetic runs:
  * every pass over an activation tensor is a kernel behind the C ABI (include/diqt.h): convolutions and their data gradients
    (`diqt_conv_*`, the data gradient of a convolution is a convolution with the flipped, transposed weights), weight gradients
    (`diqt_conv_wgrad`), GroupNorm / FiLM / Mish / SE forward (`diqt_channel_stats`, `diqt_gn_finalize`, `diqt_affine_mish`,
    `diqt_se_gate`, `diqt_scale_residual`) and reverse (`diqt_bwd_reduce`, `diqt_bwd_apply`), the loss (`diqt_loss_grad`), Adam + EMA
    (`diqt_adam_step`);
  * what is left to PyTorch is plumbing on (batch, channels)-sized vectors: the time-embedding MLP and the SE gate MLP (a few hundred
    values per sample, differentiated by autograd), the GroupNorm coefficient algebra between a reduce and an apply kernel, weight
    re-layouts, pixel (un)shuffle permutations and the channel concat of the skip connections.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import lib as L
from . import ops

def _nblk_apply(t):
    """CTAs per volume for the apply pass (it leaves no partial rows behind): eight CTAs per SM keep enough loads in flight"""
    return int(max(1, min(1184 // t.shape[0], _vox(t) // 64)))


def _nblk(t):
    """partial rows per volume for the statistics / reduce / apply kernels: about four CTAs per SM over the whole batch, at least 64 voxels each"""
    return int(max(1, min(320, 592 // t.shape[0], _vox(t) // 64)))


class _stream_ordered:
    """Inside the training step nothing reads results on the host: the per-call synchronisation of the ops wrappers is switched off, every
    kernel and every PyTorch op is ordered by the current stream (which also makes the whole step capturable in a CUDA graph)."""

    def __enter__(self):
        self.prev, ops.SYNC = ops.SYNC, False

    def __exit__(self, *exc):
        ops.SYNC = self.prev


def _dt(t):
    return L.BF16 if t.dtype == torch.bfloat16 else L.F32


def _vox(t):
    return t.shape[1] * t.shape[2] * t.shape[3]


def _unshuffle_cl(x):
    """channels-last pixel-unshuffle: (n, 2h, 2w, 2d, c) -> (n, h, w, d, 8c) with channel = c*8 + s1*4 + s2*2 + s3 (:492)."""
    n, H, W, D, c = x.shape
    v = x.reshape(n, H // 2, 2, W // 2, 2, D // 2, 2, c).permute(0, 1, 3, 5, 7, 2, 4, 6)
    return v.reshape(n, H // 2, W // 2, D // 2, c * 8).contiguous()


def _shuffle_cl(x):
    """channels-last pixel-shuffle, the inverse of `_unshuffle_cl` (PixelShuffle3D :415-438)."""
    n, h, w, d, c8 = x.shape
    c = c8 // 8
    v = x.reshape(n, h, w, d, c, 2, 2, 2).permute(0, 1, 5, 2, 6, 3, 7, 4)
    return v.reshape(n, 2 * h, 2 * w, 2 * d, c).contiguous()


def _flip_t(w):
    """weights of the data-gradient convolution: w'[ci][co][k] = w[co][ci][K-1-k]."""
    return w.detach().flip(2, 3, 4).transpose(0, 1).contiguous()


def _accum(p, g):
    g = g.to(p.dtype).reshape(p.shape)
    p.grad = g.clone() if p.grad is None else p.grad + g


def gn_backward_coefficients(S1, S2x, mean, rstd, gamma, beta, film, vox, groups):
    """The (n, c)-sized algebra between `diqt_bwd_reduce` and `diqt_bwd_apply` for y = mish(w), w = (xhat gamma + beta)(1 + scale) + shift,
    xhat = (x - mean_g) rstd_g (GroupNorm :546, FiLM :559-561).  S1 = sum_v dw, S2x = sum_v dw x with dw = dy mish'(w).
    Returns c1, c2, c3 of dx = c1 dw + c2 x + c3, d gamma, d beta and d (scale | shift) (None without FiLM); fp64 inside."""
    n, c = S1.shape
    cpg = c // groups
    S1, S2x = S1.double(), S2x.double()
    mu = mean.double().repeat_interleave(cpg, dim=1)                  # (n, c)
    r = rstd.double().repeat_interleave(cpg, dim=1)
    S2 = r * (S2x - mu * S1)                                          # sum dw * xhat
    gamma, beta = gamma.double(), beta.double()
    k = 1.0 + film[:, :c].double() if film is not None else torch.ones_like(S1)
    dbeta, dgamma = (k * S1).sum(dim=0), (k * S2).sum(dim=0)
    dfilm = torch.cat((gamma * S2 + beta * S1, S1), dim=1).float() if film is not None else None      # d scale | d shift
    cnt = float(vox * cpg)
    kg = k * gamma
    m1 = ((kg * S1).reshape(n, groups, cpg).sum(dim=2) / cnt).repeat_interleave(cpg, dim=1)
    m2 = ((kg * S2).reshape(n, groups, cpg).sum(dim=2) / cnt).repeat_interleave(cpg, dim=1)
    c1 = (r * kg).float().contiguous()
    c2 = (-r * r * m2).float().contiguous()
    c3 = (r * (r * m2 * mu - m1)).float().contiguous()
    return c1, c2, c3, dgamma, dbeta, dfilm


class UnetBackprop:
    """One training-time evaluation of `Unet.forward` and its reverse pass.  `forward` keeps what the reverse pass needs; `backward(dpred)`
    accumulates `.grad` on every parameter of the U-Net (as `loss.backward()` does in the reference)."""

    def __init__(self, unet, compute_dtype: Optional[str] = None):
        unsupported = []
        if any(m[2] is not None for m in unet.downs) or getattr(unet, "mid_attn", None) is not None:
            unsupported.append("attention blocks")
        if unet.boundary:
            unsupported.append("boundary=True")
        if any(m[0] is not None and hasattr(m[0], "deconv") for m in unet.ups):
            unsupported.append("pixel_shuffle_upsample=False")
        if not isinstance(unet.init_conv, torch.nn.Conv3d):
            unsupported.append("init_cross_embed=True")
        if unet.has_cond_image or unet.self_cond:
            unsupported.append("cond_images / self_cond")
        if unsupported:
            raise NotImplementedError("the training step does not implement: " + ", ".join(unsupported))
        self.unet = unet
        self.lib = L.load()
        mode = compute_dtype or unet.compute_dtype
        self.act = torch.bfloat16 if mode == "bf16" else torch.float32
        self.saved = None

    # ------------------------------------------------------------------ kernels
    def _stats(self, x):
        """-> per-(n, c) sums (n, c, 2) in fp64 from the partial rows of diqt_channel_stats."""
        return ops.channel_stats(x, _nblk(x)).double().sum(dim=1)

    def _gn_forward(self, x, gn, film):
        """GroupNorm (+FiLM) folded into a per-(n, c) affine, then Mish.  Returns z and what the reverse pass needs."""
        n, c = x.shape[0], x.shape[-1]
        vox, G = _vox(x), gn.num_groups
        NBLK = _nblk(x)
        part = ops.channel_stats(x, NBLK)
        a = torch.empty(n, c, dtype=torch.float32, device=x.device)
        b = torch.empty_like(a)
        gamma, beta = gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous()
        st = L.current_stream()
        L.check(self.lib.diqt_gn_finalize(part.data_ptr(), n, NBLK, vox, c, G, gn.eps, gamma.data_ptr(), beta.data_ptr(), L.ptr(film), 2 * c, 0, 1,
                                          a.data_ptr(), b.data_ptr(), st), "gn_finalize")
        z = torch.empty_like(x)
        L.check(self.lib.diqt_affine_mish(x.data_ptr(), c, z.data_ptr(), c, _dt(x), n, vox, c, a.data_ptr(), b.data_ptr(), _nblk_apply(x), 0, 0, st), "affine_mish")
        return z, dict(x=x, a=a, b=b, part=part, gn=gn, film=film)

    def _gn_backward(self, s, dz, acc=None):
        """Reverse of `_gn_forward`: returns dx (+ acc) and accumulates d gamma / d beta; FiLM gradients come back as (n, 2c) or None."""
        x, a, b, gn = s["x"], s["a"], s["b"], s["gn"]
        n, c = x.shape[0], x.shape[-1]
        vox, G = _vox(x), gn.num_groups
        cpg = c // G
        NBLK = _nblk(x)
        part = torch.empty(n, NBLK, c, 2, dtype=torch.float32, device=x.device)
        st = L.current_stream()
        L.check(self.lib.diqt_bwd_reduce(x.data_ptr(), c, dz.data_ptr(), c, _dt(x), n, vox, c, a.data_ptr(), b.data_ptr(), 1, NBLK, part.data_ptr(), st),
                "bwd_reduce")
        # the (n, c)-sized algebra (gn_backward_coefficients below states it in torch; tests/test_train_host.py checks it against autograd)
        fpart, film = s["part"], s["film"]
        c1, c2, c3 = (torch.empty(n, c, dtype=torch.float32, device=x.device) for _ in range(3))
        dgamma, dbeta = (torch.empty(n, c, dtype=torch.float32, device=x.device) for _ in range(2))     # per-volume rows
        dfilm = torch.empty(n, 2 * c, dtype=torch.float32, device=x.device) if film is not None else None
        gamma, beta = gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous()
        L.check(self.lib.diqt_gn_bwd_finalize(fpart.data_ptr(), fpart.shape[1], part.data_ptr(), NBLK, n, vox, c, G, gn.eps, gamma.data_ptr(), beta.data_ptr(),
                                              L.ptr(film), c1.data_ptr(), c2.data_ptr(), c3.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), L.ptr(dfilm), st),
                "gn_bwd_finalize")
        _accum(gn.bias, dbeta.sum(dim=0))
        _accum(gn.weight, dgamma.sum(dim=0))
        dx = torch.empty_like(x)
        L.check(self.lib.diqt_bwd_apply(x.data_ptr(), c, dz.data_ptr(), c, L.ptr(acc), c, dx.data_ptr(), c, _dt(x), n, vox, c, a.data_ptr(), b.data_ptr(),
                                        c1.data_ptr(), c2.data_ptr(), c3.data_ptr(), 1, _nblk_apply(x), st), "bwd_apply")
        return dx, dfilm

    def _conv(self, x, conv, mode="k3"):
        return ops.conv3d(x, conv.weight, conv.bias, mode=mode)

    def _conv_backward(self, x, conv, dy, mode="k3", need_dx=True):
        """Weight / bias gradients of a 3x3x3 or 1x1x1 convolution and (optionally) its data gradient."""
        c_in, c_out = x.shape[-1], dy.shape[-1]
        taps = 27 if mode == "k3" else 1
        _accum(conv.weight, self._wgrad(x, dy, taps))
        if conv.bias is not None:
            _accum(conv.bias, self._stats(dy)[..., 0].sum(dim=0).float())
        if not need_dx:
            return None
        return ops.conv3d(dy, _flip_t(conv.weight), None, mode=mode)

    def _wgrad(self, x, dy, taps, c_in=None):
        n, d0, d1, d2, ld_x = x.shape
        c_in = ld_x if c_in is None else c_in
        c_out = dy.shape[-1]
        nbytes = C.c_size_t(0)
        L.check(self.lib.diqt_conv_wgrad_workspace_bytes(n, d0, d1, d2, c_in, c_out, taps, C.byref(nbytes)), "conv_wgrad_workspace_bytes")
        ws = torch.empty(nbytes.value // 4, dtype=torch.float32, device=x.device)
        dw = torch.empty(c_out, c_in, taps, dtype=torch.float32, device=x.device)
        L.check(self.lib.diqt_conv_wgrad(x.data_ptr(), ld_x, dy.data_ptr(), c_out, _dt(x), n, d0, d1, d2, c_in, c_out, taps, L.IMPL_AUTO, dw.data_ptr(),
                                         ws.data_ptr(), L.current_stream()), "conv_wgrad")
        return dw

    def _add(self, a, b):
        return ops.rows_combine(a.reshape(-1, a.shape[-1]), 0, b.reshape(-1, b.shape[-1])).reshape(a.shape)

    # ------------------------------------------------------------------ ResnetBlock :568-614
    def _resnet_forward(self, blk, x, t_act):
        c_out = blk.dim_out
        te = F.linear(t_act, blk.time_mlp[1].weight, blk.time_mlp[1].bias)      # (n, 2 c_out), autograd keeps the graph
        film = te.detach().float().contiguous()
        z1, s1 = self._gn_forward(x, blk.block1.groupnorm, None)
        h1 = self._conv(z1, blk.block1.project)
        z2, s2 = self._gn_forward(h1, blk.block2.groupnorm, film)
        h2 = self._conv(z2, blk.block2.project)
        res = self._conv(x, blk.res_conv, "k1") if blk.has_res_conv else x
        sv = dict(blk=blk, x=x, z1=z1, z2=z2, h2=h2, s1=s1, s2=s2, te=te)
        if blk.has_se:
            w1, w2 = blk.se.fc[0].weight, blk.se.fc[2].weight
            out, gate, _, hpart = ops.se_scale_residual(h2, res, w1, w2, _nblk(h2), return_stats=True)
            sv["gate"], sv["hpart"] = gate, hpart
        else:
            out = self._add(h2, res)
        return out, sv

    def _resnet_backward(self, sv, d_out):
        blk, x, h2 = sv["blk"], sv["x"], sv["h2"]
        n, c = h2.shape[0], h2.shape[-1]
        vox = _vox(h2)
        NBLK = _nblk(h2)
        st = L.current_stream()
        # residual branch
        if blk.has_res_conv:
            dx_res = self._conv_backward(x, blk.res_conv, d_out, "k1")
        else:
            dx_res = d_out
        # SE gate: h3 = h2 * gate(mean(h2))
        if blk.has_se:
            part = torch.empty(n, NBLK, c, 2, dtype=torch.float32, device=h2.device)
            L.check(self.lib.diqt_bwd_reduce(h2.data_ptr(), c, d_out.data_ptr(), c, _dt(h2), n, vox, c, 0, 0, 0, NBLK, part.data_ptr(), st), "bwd_reduce")
            w1 = blk.se.fc[0].weight.detach().float().contiguous()
            w2 = blk.se.fc[2].weight.detach().float().contiguous()
            hpart, gate = sv["hpart"], sv["gate"].contiguous()
            hid = w1.shape[0]
            c3 = torch.empty(n, c, dtype=torch.float32, device=h2.device)
            dw1, dw2 = torch.empty(n, *w1.shape, dtype=torch.float32, device=h2.device), torch.empty(n, *w2.shape, dtype=torch.float32, device=h2.device)
            L.check(self.lib.diqt_se_bwd(hpart.data_ptr(), hpart.shape[1], part.data_ptr(), NBLK, n, vox, c, hid, w1.data_ptr(), w2.data_ptr(), gate.data_ptr(),
                                         c3.data_ptr(), dw1.data_ptr(), dw2.data_ptr(), st), "se_bwd")
            _accum(blk.se.fc[0].weight, dw1.sum(dim=0))
            _accum(blk.se.fc[2].weight, dw2.sum(dim=0))
            c1 = gate
            d_h2 = torch.empty_like(h2)
            L.check(self.lib.diqt_bwd_apply(0, c, d_out.data_ptr(), c, 0, c, d_h2.data_ptr(), c, _dt(h2), n, vox, c, 0, 0, c1.data_ptr(), 0, c3.data_ptr(), 0,
                                            _nblk_apply(h2), st), "bwd_apply")
        else:
            d_h2 = d_out
        d_z2 = self._conv_backward(sv["z2"], blk.block2.project, d_h2)
        d_h1, dfilm = self._gn_backward(sv["s2"], d_z2)
        d_z1 = self._conv_backward(sv["z1"], blk.block1.project, d_h1)
        dx, _ = self._gn_backward(sv["s1"], d_z1, acc=dx_res)
        self._film_grads.append((sv["te"], dfilm))
        return dx

    # ------------------------------------------------------------------ Unet.forward :1554-1684
    def forward(self, x, time, lowres_cond_img=None):
        with _stream_ordered():
            return self._forward(x, time, lowres_cond_img)

    def backward(self, dpred):
        """dpred: (n, c_out, S, S, S) fp32 gradient of the loss with respect to the prediction."""
        with _stream_ordered():
            return self._backward(dpred)

    def _forward(self, x, time, lowres_cond_img=None):
        unet = self.unet
        assert not (unet.lowres_cond and lowres_cond_img is None), 'low resolution conditioning image must be present'
        dev = x.device
        if not x.is_cuda:
            raise RuntimeError("the training step runs only on a CUDA device (sm_100a kernels; there is no CPU fallback)")
        self._film_grads = []
        tape: List = []
        inp = torch.cat((x, lowres_cond_img), dim=1) if lowres_cond_img is not None else x          # :1576
        n, cin = inp.shape[:2]
        # init_conv :1291: channels zero-padded to 16 so that the general convolution kernels apply
        # (bf16: padded to 64 so that the conv and its weight gradient run on the tensor cores: 50 + 75 us instead of 460 + 1080 us on the
        # CUDA-core kernels at 64^3, profiles/launch_summary_train_r3h.txt; fp32 exact mode: 16, the CUDA-core kernels' granularity)
        cpad = 64 if self.act == torch.bfloat16 and unet.init_conv.weight.shape[0] % 64 == 0 else 16
        x16 = torch.zeros(n, *inp.shape[2:], cpad, dtype=self.act, device=dev)
        x16[..., :cin] = inp.permute(0, 2, 3, 4, 1).to(self.act)
        w = unet.init_conv.weight
        w16 = torch.zeros(w.shape[0], cpad, *w.shape[2:], dtype=w.dtype, device=dev)
        w16[:, :cin] = w.detach()
        h = ops.conv3d(x16, w16, unet.init_conv.bias, mode="k3")
        tape.append(("init", x16, cin))
        # time embedding :518-533, 1305-1316 (a (n, 4 dim) vector: PyTorch, autograd)
        pe = unet.to_time_hiddens[0].weights
        tcol = time[:, None].to(torch.float32)
        freqs = tcol * pe[None, :] * 2 * math.pi
        four = torch.cat((tcol, freqs.sin(), freqs.cos()), dim=-1)
        hid = F.mish(F.linear(four, unet.to_time_hiddens[1].weight, unet.to_time_hiddens[1].bias))
        t = F.linear(hid, unet.to_time_cond[0].weight, unet.to_time_cond[0].bias)
        t_act = F.mish(t)                                           # ResnetBlock.time_mlp[0] (:586)
        nl = len(unet.downs)
        hiddens = []
        for l, (_, init_block, _attn, blocks, post) in enumerate(unet.downs):
            h, sv = self._resnet_forward(init_block, h, t_act)
            tape.append(("res", sv))
            for blk in blocks:
                h, sv = self._resnet_forward(blk, h, t_act)
                tape.append(("res", sv))
            if l != nl - 1:
                hiddens.append(h)
                tape.append(("down", h, post[1]))
                h = ops.conv3d(h, post[1].weight, post[1].bias, mode="down")
            else:
                tape.append(("k1", h, post))
                h = self._conv(h, post, "k1")
        if unet.deep_feature:
            h, sv = self._resnet_forward(unet.mid_block, h, t_act)
            tape.append(("res", sv))
        for u, (up, first, blocks) in enumerate(unet.ups):
            if up is not None:
                conv = up.net[0]
                pre = self._conv(h, conv, "k1")
                tape.append(("up", h, pre, conv))
                n_, c8 = pre.shape[0], pre.shape[-1]
                one = torch.ones(n_, c8, dtype=torch.float32, device=dev)
                zero = torch.zeros_like(one)
                act = torch.empty_like(pre)
                L.check(self.lib.diqt_affine_mish(pre.data_ptr(), c8, act.data_ptr(), c8, _dt(pre), n_, _vox(pre), c8, one.data_ptr(), zero.data_ptr(), _nblk_apply(pre), 0, 0,
                                                  L.current_stream()), "affine_mish")
                h = _shuffle_cl(act)
                skip = hiddens.pop()
                if unet.skip_connect_scale != 1.:
                    skip = skip * unet.skip_connect_scale
                tape.append(("cat", h.shape[-1]))
                h = torch.cat((h, skip), dim=-1)                    # :1653
            h, sv = self._resnet_forward(first, h, t_act)
            tape.append(("res", sv))
            for blk in blocks:
                h, sv = self._resnet_forward(blk, h, t_act)
                tape.append(("res", sv))
        if unet.final_res_block is not None:
            h, sv = self._resnet_forward(unet.final_res_block, h, t_act)
            tape.append(("res", sv))
        # final_conv :1477 in fp32 (one output channel, padded to four for the general kernel)
        fc = unet.final_conv
        k = fc.weight.shape[-1]
        hf = h.float()
        co = fc.weight.shape[0]
        w4 = torch.zeros(4 * ((co + 3) // 4), *fc.weight.shape[1:], dtype=torch.float32, device=dev)
        w4[:co] = fc.weight.detach()
        b4 = torch.zeros(w4.shape[0], dtype=torch.float32, device=dev)
        b4[:co] = fc.bias.detach()
        pred = ops.conv3d(hf, w4, b4, mode="k3" if k == 3 else "k1")[..., :co]
        tape.append(("final", hf, h.dtype, co, k))
        self.saved = dict(tape=tape, skip_grads={})
        return pred.permute(0, 4, 1, 2, 3).contiguous()

    def _backward(self, dpred):
        assert self.saved is not None, "backward() needs a forward() first"
        unet, tape = self.unet, self.saved["tape"]
        dev = dpred.device
        d = None
        skip_grads: List = []
        for rec in reversed(tape):
            kind = rec[0]
            if kind == "final":
                _, hf, act_dtype, co, k = rec
                fc = unet.final_conv
                taps = 27 if k == 3 else 1
                dy = dpred.permute(0, 2, 3, 4, 1).contiguous().float()                               # (n, S, S, S, co)
                if k == 1 and co <= 16:
                    # one output channel: the roles swapped (dw[co][ci] = sum_v dy[v][co] hf[v][ci] is the narrow-input gradient of a conv
                    # that maps dy to hf), so the narrow kernel applies instead of a 64 x 64 tile with one live row
                    _accum(fc.weight, self._wgrad(dy, hf, 1).transpose(0, 1).contiguous())
                else:
                    _accum(fc.weight, self._wgrad(hf, dy, taps))
                _accum(fc.bias, dy.sum(dim=(0, 1, 2, 3)))
                dy16 = torch.zeros(*dy.shape[:-1], 16, dtype=torch.float32, device=dev)
                dy16[..., :co] = dy
                wt = torch.zeros(16, *fc.weight.shape[1:], dtype=torch.float32, device=dev)
                wt[:co] = fc.weight.detach()
                d = ops.conv3d(dy16, _flip_t(wt), None, mode="k3" if k == 3 else "k1").to(act_dtype)
            elif kind == "res":
                d = self._resnet_backward(rec[1], d)
            elif kind == "cat":
                c_up = rec[1]
                skip = d[..., c_up:].contiguous()
                if unet.skip_connect_scale != 1.:
                    skip = skip * unet.skip_connect_scale
                skip_grads.append(skip)
                d = d[..., :c_up].contiguous()
            elif kind == "up":
                _, h_in, pre, conv = rec
                g = _unshuffle_cl(d)
                n_, c8 = pre.shape[0], pre.shape[-1]
                one = torch.ones(n_, c8, dtype=torch.float32, device=dev)
                zero = torch.zeros_like(one)
                d_pre = torch.empty_like(pre)
                L.check(self.lib.diqt_bwd_apply(pre.data_ptr(), c8, g.data_ptr(), c8, 0, c8, d_pre.data_ptr(), c8, _dt(pre), n_, _vox(pre), c8, one.data_ptr(),
                                                zero.data_ptr(), one.data_ptr(), 0, 0, 1, _nblk_apply(pre), L.current_stream()), "bwd_apply")
                d = self._conv_backward(h_in, conv, d_pre, "k1")
            elif kind == "k1":
                _, h_in, conv = rec
                d = self._conv_backward(h_in, conv, d, "k1")
            elif kind == "down":
                _, h_in, conv = rec
                xu = _unshuffle_cl(h_in)
                dxu = self._conv_backward(xu, conv, d, "k1")
                d = self._add(_shuffle_cl(dxu), self._pop_skip(skip_grads))
            elif kind == "init":
                _, x16, cin = rec
                conv = unet.init_conv
                if x16.shape[-1] == 64:      # tensor-core weight gradient over the padded channels, the real ones kept
                    _accum(conv.weight, self._wgrad(x16, d, 27).reshape(conv.weight.shape[0], 64, *conv.weight.shape[2:])[:, :cin])
                else:
                    _accum(conv.weight, self._wgrad(x16, d, 27, c_in=cin))      # only the real input channels (row pitch 16)
                _accum(conv.bias, self._stats(d)[..., 0].sum(dim=0).float())
        # FiLM rows -> time MLPs (autograd on (n, 2c) vectors)
        tes = [te for te, g in self._film_grads if g is not None]
        gs = [g.to(te.dtype) for te, g in self._film_grads if g is not None]
        if tes:
            torch.autograd.backward(tes, gs)
        self.saved = None

    @staticmethod
    def _pop_skip(skip_grads):
        # the decoder pops the encoder's hiddens last-in first-out (:1653), so walking the tape backwards meets the encoder levels in the
        # order their skip gradients were produced last: the most recently produced gradient belongs to the shallowest level
        return skip_grads.pop()


class _LossBackward(torch.autograd.Function):
    """Lets `loss.backward()` (ImagenTrainer.forward -> accelerator.backward, trainer.py:1124) start the hand-written reverse pass."""

    @staticmethod
    def forward(ctx, anchor, loss_value, runner, dpred):
        ctx.runner, ctx.dpred = runner, dpred
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.runner.backward(ctx.dpred * grad_out.to(ctx.dpred.dtype))
        return None, None, None, None


_LOSS_KIND = {"l1": 0, "l2": 1, "huber": 2}


def p_losses(imagen, unet, x_start, times, *, noise_scheduler, lowres_cond_img=None, noise=None, pred_objective="noise", p2_loss_weight_gamma=0.,
             compute_dtype=None):
    """imagen_pytorch3D.py:2277-2373 on the kernels: q_sample, U-Net forward with the tape, loss and d loss / d pred.
    Returns (loss, pred, x_noisy, lowres_cond_img) like the reference; `loss.backward()` runs `UnetBackprop.backward`."""
    lib = L.load()
    noise = torch.randn_like(x_start) if noise is None else noise
    x_start = imagen.normalize_img(x_start)
    lowres_cond_img = imagen.normalize_img(lowres_cond_img) if lowres_cond_img is not None else None
    x_noisy, log_snr, alpha, sigma = noise_scheduler.q_sample(x_start=x_start, t=times, noise=noise)
    noise_cond = noise_scheduler.get_condition(times)
    runner = UnetBackprop(unet, compute_dtype)
    pred = runner.forward(x_noisy.float(), noise_cond, lowres_cond_img=lowres_cond_img)
    if pred_objective == "noise":
        target = noise
    elif pred_objective == "x_start":
        target = x_start
    elif pred_objective == "v":
        target = alpha * noise - sigma * x_start
    else:
        raise ValueError(f"unknown objective {pred_objective}")
    n = pred.shape[0]
    count = pred.numel() // n
    weight = torch.ones(n, dtype=torch.float32, device=pred.device)
    if p2_loss_weight_gamma > 0:
        weight = (imagen.p2_loss_weight_k + log_snr.exp()) ** -p2_loss_weight_gamma
        weight = weight.reshape(n).float()
    sw = (weight / float(n * count)).contiguous()
    target = target.float().contiguous()
    dpred = torch.empty_like(pred)
    nblk = 64
    part = torch.empty(n, nblk, dtype=torch.float32, device=pred.device)
    clamp = 1 if pred_objective == "x_start" else 0
    L.check(lib.diqt_loss_grad(pred.data_ptr(), target.data_ptr(), n, count, _LOSS_KIND[imagen.loss_type], clamp, float(imagen.min_bound), sw.data_ptr(),
                               dpred.data_ptr(), part.data_ptr(), nblk, L.current_stream()), "loss_grad")
    if clamp:
        pred = pred.clamp(min=imagen.min_bound)                    # the reference clamps pred in place (:2353) and returns it
    losses = part.double().sum(dim=1).float() / count               # per-sample means
    loss_value = (losses * weight).mean()
    anchor = torch.zeros((), device=pred.device, requires_grad=True)
    loss = _LossBackward.apply(anchor, loss_value, runner, dpred)
    return loss, pred, x_noisy, lowres_cond_img


def allreduce_gradients(params, group=None):
    """Data-parallel training, one process per GPU (the reference wraps the U-Net in DistributedDataParallel through accelerate,
    trainer.py:476-499): the gradients of all ranks are averaged with ONE all-reduce over a flat buffer (NCCL over NVLink on the GPU box,
    gloo in the CPU tests) before the optimizer step.  Parameters without a gradient on this rank contribute zeros."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    params = [p for p in params]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        p.grad = flat[off:off + n].reshape(p.shape).to(p.dtype)
        off += n


class AdamState:
    """torch.optim.Adam's arithmetic (trainer.py: `Adam(unet.parameters(), lr, eps, betas)`) as one kernel per parameter tensor, with the
    exponential moving average of ImagenTrainer.update folded into the same pass."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.m = [torch.zeros_like(p, dtype=torch.float32) for p in self.params]
        self.v = [torch.zeros_like(p, dtype=torch.float32) for p in self.params]
        self.steps = 0

    @torch.no_grad()
    def step(self, ema_params=None, ema_decay=0.0, grad_scale=1.0):
        lib = L.load()
        self.steps += 1
        st = L.current_stream()
        for i, p in enumerate(self.params):
            if p.grad is None:
                continue
            g = p.grad.detach().float().contiguous()
            ema = ema_params[i] if ema_params is not None else None
            assert p.is_contiguous() and p.dtype == torch.float32
            L.check(lib.diqt_adam_step(p.data_ptr(), g.data_ptr(), self.m[i].data_ptr(), self.v[i].data_ptr(), p.numel(), self.lr, self.betas[0],
                                       self.betas[1], self.eps, self.weight_decay, self.steps, grad_scale, L.ptr(ema), ema_decay, st), "adam_step")
            p._version  # noqa: B018  (parameters were written behind autograd's back: callers refresh the engines, see Unet.refresh_weights)

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    # ---- the layout of torch.optim.Adam.state_dict() (what trainer.py:858 stores under 'optim{i}'), so that checkpoints written by the
    # reference resume here and the other way round: state[i] = {step, exp_avg, exp_avg_sq} by parameter index, one param group
    def state_dict(self):
        state = {}
        if self.steps > 0:
            for i in range(len(self.params)):
                state[i] = dict(step=torch.tensor(float(self.steps)), exp_avg=self.m[i].detach().clone(), exp_avg_sq=self.v[i].detach().clone())
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False, maximize=False, foreach=None,
                     capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False, params=list(range(len(self.params))))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError(f"optimizer state has {len(order)} parameters, this U-Net {len(self.params)}")
        g0 = groups[0]
        self.lr, self.betas, self.eps = g0["lr"], tuple(g0["betas"]), g0["eps"]
        self.weight_decay = g0.get("weight_decay", 0.0)
        steps = 0
        for pos, idx in enumerate(order):
            st = sd["state"].get(idx)
            if st is None:
                self.m[pos].zero_(); self.v[pos].zero_()
                continue
            self.m[pos].copy_(st["exp_avg"].to(self.m[pos].device, torch.float32))
            self.v[pos].copy_(st["exp_avg_sq"].to(self.v[pos].device, torch.float32))
            steps = max(steps, int(float(st["step"])))
        self.steps = steps
