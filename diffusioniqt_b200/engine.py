"""UnetEngine: one compiled instance of the 3-D U-Net forward for a fixed (batch, volume, dtype).

It owns the device buffers (channels-last activations, packed weights, statistics scratch), the
conv plans (TMA descriptors) and an ordered list of kernel launches that reproduces
`Unet.forward` (/root/reference/imagen_pytorch3D.py:1554-1684).  PyTorch is used for device
memory and streams only; every arithmetic op is a call into libdiqt_b200.so.

Fusions relative to the reference's op list:
  * GroupNorm + FiLM + Mish  -> channel-stats pass, tiny finalize, one affine+Mish pass
  * SE pool / gate / scale + residual add -> stats pass, tiny gate kernel, one fused pass that also
    emits the statistics the next block's GroupNorm needs
  * skip concat (:1653)      -> producers write straight into a row-pitched concat buffer
  * pixel-unshuffle / pixel-shuffle (+Mish) -> folded into the 1x1x1 conv's loads / stores
  * time MLPs of all blocks  -> one dense layer, evaluated once per sampler for all steps
  * final 1x1x1 conv + clamp + posterior + noise (:1976-2056) -> one kernel
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional

import torch

from . import lib as L


class Act:
    """A channels-last activation view: n volumes x voxels rows of `c` channels, row pitch `ld`."""

    __slots__ = ("buf", "ptr", "c", "ld", "off", "stats")

    def __init__(self, buf: torch.Tensor, c: int, ld: int, offset: int = 0):
        self.buf = buf
        self.ptr = buf.data_ptr() + offset * buf.element_size()
        self.c, self.ld, self.off = c, ld, offset
        self.stats = None  # (partial tensor, nblk) when a producer already reduced this tensor


def _nblk(n: int, voxels: int, per_device: int = 148) -> int:
    """Blocks per volume: reductions use ~one 512-thread block per SM (few partial sums to finalize),
    pure streaming kernels ~four 256-thread blocks per SM."""
    return max(1, min(voxels // 128, max(1, per_device // n)))


def deconv_as_conv3_weights(w, b):
    """nn.ConvTranspose3d(k 3, stride 2, padding 1, output_padding 1) (imagen_pytorch3D.py:445-447) as a 3x3x3 conv (padding 1) to 8 c_out
    channels followed by a pixel shuffle: output voxel 2 j + p takes, per axis, kernel tap 1 of input j (p = 0), or tap 2 of input j and tap 0
    of input j + 1 (p = 1).  w: (c_in, c_out, 3, 3, 3) -> (8 c_out, c_in, 3, 3, 3) with channel = c_out_index * 8 + p1 * 4 + p2 * 2 + p3
    (the order PixelShuffle3D :415-438 reads them in); the unused taps are zero."""
    ci, co = w.shape[:2]
    kmap = {(0, 1): 1, (1, 1): 2, (1, 2): 0}                      # (parity, conv tap a = offset + 1) -> transposed-conv tap
    wc = torch.zeros(co, 2, 2, 2, ci, 3, 3, 3, dtype=w.dtype, device=w.device)
    for p1 in range(2):
        for p2 in range(2):
            for p3 in range(2):
                for a in range(3):
                    for b_ in range(3):
                        for c in range(3):
                            ks = (kmap.get((p1, a)), kmap.get((p2, b_)), kmap.get((p3, c)))
                            if None not in ks:
                                wc[:, p1, p2, p3, :, a, b_, c] = w[:, :, ks[0], ks[1], ks[2]].t()
    return wc.reshape(co * 8, ci, 3, 3, 3), b.repeat_interleave(8)


class UnetEngine:
    def __init__(self, unet, batch: int, dims, dtype: str = "bf16", device=None, conv_impl: str = "auto", taps=()):
        self.lib = L.load()
        # names of ResnetBlocks (the oracle's tap names, e.g. "downs.0.1") whose output is copied aside during the forward: parity tests
        # compare intermediate activations at full size without a second code path
        self.tap_names = tuple(taps)
        self.taps = {}
        if not torch.cuda.is_available():
            raise L.DiqtError("UnetEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        self.unet = unet
        self.n = int(batch)
        self.dims0 = tuple(int(d) for d in dims)
        assert len(self.dims0) == 3
        self.dtype = dtype
        self.tdtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        self.ddtype = L.BF16 if dtype == "bf16" else L.F32
        self.impl = {"auto": L.IMPL_AUTO, "simt": L.IMPL_SIMT, "tc": L.IMPL_TC, "zm": L.IMPL_ZM}[conv_impl]
        if dtype == "fp32":
            self.impl = L.IMPL_SIMT
        # boundary mode (imagen_pytorch3D.py:37-46): the f^3 sub-volumes are stored merged as one (f*h)^3 volume, so every conv is a
        # plain zero-padded conv over the merged volume (== boundary_pad + unpadded conv) while statistics stay per sub-volume
        self.sub_f = int(unet.batch_sample_factor) if unet.boundary else 0
        if self.sub_f > 1:
            if self.n != self.sub_f ** 3 or len(set(self.dims0)) != 1:
                raise ValueError(f"boundary mode needs batch == batch_sample_factor^3 = {self.sub_f ** 3} cubic sub-volumes, got batch {self.n}, dims {self.dims0}")
        nl = len(unet.in_out)
        for d in self.dims0:
            if d % (1 << (nl - 1)) != 0:
                raise ValueError(f"volume side {d} is not divisible by 2^{nl - 1}")
        self.nl = nl
        self.level_dims = [tuple(d >> l for d in self.dims0) for l in range(nl)]
        self.level_vox = [a * b * c for (a, b, c) in self.level_dims]          # rows per statistics volume
        if self.sub_f > 1:
            self.conv_n = 1
            self.conv_dims = [tuple(self.sub_f * d for d in ld) for ld in self.level_dims]
        else:
            self.conv_n, self.conv_dims = self.n, self.level_dims
        self.level_sub = [(self.sub_f, ld[0]) if self.sub_f > 1 else (0, 0) for ld in self.level_dims]
        self._plans: List[int] = []
        self._keep: List[torch.Tensor] = []   # packed weights etc.
        self._ops = []                        # callables(stream)
        self._film_rows = 0
        self.film_row_ptr = 0                 # device int32* (sampler step) or NULL
        self.film_stride_n = 1
        self.conv_impls = {}                  # site name -> resolved impl (for tests / reporting)
        self.fused_gn = []                    # conv sites that apply GroupNorm + FiLM + Mish on their own load path
        self.split_k = []                     # conv sites of the per-tap family that run with split-K (small volumes)
        # sampler states (schedule tables, captured step graphs) that bake this engine's buffer addresses live and die with it
        self.sampler_cache = {}
        self.film_gen = 0
        self._attn_plans: List[int] = []
        self._linattn_plans: List[int] = []
        self.attn_impls = {}                  # attention site -> "tc" | "simt"
        self._build()

    # ------------------------------------------------------------------ helpers
    def _empty(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=dtype or self.tdtype, device=self.device)

    def _f32(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self._keep.append(t)
        return t

    def _conv_desc(self, mode, level_in, c_in, ld_in, c_out, ld_out, vol=None, impl=None):
        cn, (d0, d1, d2) = (self.conv_n, self.conv_dims[level_in]) if vol is None else (vol[0], vol[1:])
        return L.ConvDesc(mode=mode, dtype=self.ddtype, impl=self.impl if impl is None else impl, n=cn, d0=d0, d1=d1, d2=d2,
                          c_in=c_in, ld_in=ld_in, c_out=c_out, ld_out=ld_out, flags=0)

    def _resolves_to_zm(self, desc) -> bool:
        impl = C.c_int(0)
        return self.lib.diqt_conv_resolved_impl(C.byref(desc), C.byref(impl)) == 0 and impl.value == L.IMPL_ZM

    def _conv_site(self, name, mode, level_in, c_in, ld_in, c_out, ld_out, weight, bias, src_ptr, dst_ptr, stats=None, vol=None, impl=None, gn=None):
        """Pack weights, create the plan, return the launch closure.  With `stats` (index of a statistics scratch set) the conv
        is asked to emit the channel statistics of its output; if it cannot, a separate stats pass is appended by the caller.
        `vol` = (n, d0, d1, d2) overrides the level geometry (token volumes of the attention blocks).
        `gn` = (grouped stats record of the input, nn.GroupNorm holder, FiLM column offset or None): the conv reads the RAW input and
        applies GroupNorm + FiLM + Mish on its own load path (include/diqt.h: diqt_conv_plan_set_gn); `gn` = "affine": the same with
        the per-(volume, channel) affine that diqt_gn_finalize leaves in self.aff_a / self.aff_b (diqt_conv_plan_set_gn_affine)."""
        cn, (d0, d1, d2) = (self.conv_n, self.conv_dims[level_in]) if vol is None else (vol[0], vol[1:])
        if self.sub_f > 1:
            stats = None   # fused conv statistics are per conv volume; boundary mode needs them per sub-volume
        desc = L.ConvDesc(mode=mode, dtype=self.ddtype, impl=self.impl if impl is None else impl, n=cn, d0=d0, d1=d1, d2=d2,
                          c_in=c_in, ld_in=ld_in, c_out=c_out, ld_out=ld_out, flags=0)
        impl = C.c_int(0)
        L.check(self.lib.diqt_conv_resolved_impl(C.byref(desc), C.byref(impl)), f"conv {name}")
        self.conv_impls[name] = impl.value
        nbytes = C.c_size_t(0)
        L.check(self.lib.diqt_conv_packed_bytes(C.byref(desc), C.byref(nbytes)), f"conv {name}")
        packed = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        pbias = torch.empty(c_out, dtype=torch.float32, device=self.device)
        w = weight.detach().to(device=self.device, dtype=torch.float32)
        if self.dtype == "bf16":
            w = w.to(torch.bfloat16).to(torch.float32)  # both kernel families see the same bf16-rounded weights
        w = w.contiguous()
        b = bias.detach().to(device=self.device, dtype=torch.float32).contiguous() if bias is not None else None
        st = L.current_stream()
        L.check(self.lib.diqt_conv_pack(C.byref(desc), w.data_ptr(), L.ptr(b), packed.data_ptr(), pbias.data_ptr(), st), f"pack {name}")
        torch.cuda.current_stream().synchronize()  # w / b temporaries may be freed after this
        self._keep += [packed, pbias]
        plan = C.c_void_p(0)
        L.check(self.lib.diqt_conv_plan_create(C.byref(desc), src_ptr, dst_ptr, packed.data_ptr(), pbias.data_ptr(), C.byref(plan)),
                f"plan {name}")
        self._plans.append(plan.value)
        run, pv = self.lib.diqt_conv_run, plan.value
        # small volumes: the per-tap family splits K over several CTAs and wants a scratch buffer for the fp32 partial tiles
        wsb = C.c_size_t(0)
        L.check(self.lib.diqt_conv_plan_workspace_bytes(pv, C.byref(wsb)), f"workspace_bytes {name}")
        if wsb.value and os.environ.get("DIQT_DISABLE_SPLITK", "0") != "1":
            ws = torch.zeros(wsb.value, dtype=torch.uint8, device=self.device)
            self._keep.append(ws)
            L.check(self.lib.diqt_conv_plan_set_workspace(pv, ws.data_ptr(), wsb.value), f"set_workspace {name}")
            self.split_k.append(name)
        film_off = None
        if gn == "affine":
            L.check(self.lib.diqt_conv_plan_set_gn_affine(pv, self.aff_a.data_ptr(), self.aff_b.data_ptr()), f"set_gn_affine {name}")
            self.fused_gn.append(name)
        elif gn is not None:
            (gpart, gnb, ggrp, gng), gmod, film_off = gn
            gamma, beta = self._f32(gmod.weight), self._f32(gmod.bias)
            L.check(self.lib.diqt_conv_plan_set_gn(pv, ggrp.data_ptr(), gng, self.level_vox[level_in], gmod.num_groups, float(gmod.eps),
                                                   gamma.data_ptr(), beta.data_ptr()), f"set_gn {name}")
            self.fused_gn.append(name)
        self._last_conv_stats = None          # (partial, nblk, group, ngroups) when the conv emits the statistics of its output
        if stats is not None:
            nb, ng = C.c_int(0), C.c_int(0)
            part = self.part[stats]
            if self.grouped:
                L.check(self.lib.diqt_conv_plan_set_stats_g(pv, part.data_ptr(), self.grp[stats].data_ptr(), self.tick[stats].data_ptr(),
                                                            C.byref(nb), C.byref(ng)), f"set_stats_g {name}")
            else:
                L.check(self.lib.diqt_conv_plan_set_stats(pv, part.data_ptr(), C.byref(nb)), f"set_stats {name}")
            if nb.value:
                self._last_conv_stats = (part, nb.value, self.grp[stats] if ng.value else None, ng.value)
        if film_off is not None:
            eng, set_film = self, self.lib.diqt_conv_plan_set_film

            def op(st):          # the FiLM table may move between calls (set_condition): bind its address at launch / capture time
                L.check(set_film(pv, eng.film.data_ptr() + film_off * 4, eng._film_cols, eng.film_row_ptr, eng.film_stride_n), name)
                L.check(run(pv, st), name)
            return op
        return lambda st: L.check(run(pv, st), name)

    # ------------------------------------------------------------------ build
    def _build(self):
        u, lib, n = self.unet, self.lib, self.n
        nl, dims = self.nl, u.dims
        nblocks = u.num_resnet_blocks
        ops = self._ops
        dd = self.ddtype

        # ---- inputs (static, so that a captured graph can be replayed)
        ch = u.channels
        sp = self.dims0
        self.x_in = self._empty(n, ch, *sp, dtype=torch.float32)
        self.lowres = self._empty(n, ch, *sp, dtype=torch.float32) if u.lowres_cond else None
        self.cond_images = self._empty(n, u.cond_images_channels, *sp, dtype=torch.float32) if u.has_cond_image else None
        vox0 = self.level_vox[0]
        planes, strides = [], []
        # channel order of the reference's concats (:1569-1584): [cond_images, x, lowres]
        for t in (self.cond_images, self.x_in, self.lowres):
            if t is None:
                continue
            for c in range(t.shape[1]):
                planes.append(t.data_ptr() + c * vox0 * 4)
                strides.append(t.shape[1] * vox0)
        assert len(planes) == u.init_channels, (len(planes), u.init_channels)
        if len(planes) > 8:
            raise NotImplementedError("init_conv with more than 8 input channels")
        self._planes = (C.c_void_p * len(planes))(*planes)
        self._pstrides = (C.c_int64 * len(planes))(*strides)

        # ---- widest channel count at each resolution level (sizes the shared scratch buffers)
        need = [0] * nl
        for l in range(nl):
            need[l] = max(need[l], dims[l], dims[l + 1] if l == nl - 1 else 0)
        for l in range(nl - 1):
            need[l] = max(need[l], dims[l + 1] + dims[l])
        need[0] = max(need[0], u._locals["dim"], dims[0])
        cmax = max(need)

        # ---- scratch for statistics / affine parameters
        self.nblk = [_nblk(n, v) for v in self.level_vox]
        self.nblk_stream = [_nblk(n, v, 592) for v in self.level_vox]
        pmax = 304 * n * cmax * 2     # up to two partials (z-march slots) per SM and volume
        # statistics scratch sets: 0 / 1 ping-pong between block outputs, 2 = conv1 output, 3 = conv2 output (a conv with the fused
        # input GroupNorm READS its input's set in its prologue and WRITES its output's set at its end: they must not be the same set)
        self.part = [torch.zeros(pmax, dtype=torch.float32, device=self.device) for _ in range(4)]
        # grouped statistics (include/diqt.h): producers also reduce their partial rows in <= 16 groups and the consumers finalise
        # GroupNorm / SE in their own prologue, which removes ~57 single-CTA finalize launches per forward.  Plain small batches only.
        self.grouped = self.sub_f <= 1 and n <= 2 and os.environ.get("DIQT_DISABLE_GROUPED", "0") != "1"   # variable: A/B and diagnostics
        self.grp = [torch.zeros(16 * n * cmax * 2, dtype=torch.float32, device=self.device) for _ in range(4)]
        self.tick = [torch.zeros(16 * n, dtype=torch.int32, device=self.device) for _ in range(4)]
        self.gn_fusion_min = int(os.environ.get("DIQT_GN_FUSION_MIN", str(1 << 21)))   # voxels x input channels; variable: A/B, tests
        self.fuse_gn = self.grouped and self.dtype == "bf16" and os.environ.get("DIQT_DISABLE_GN_FUSION", "0") != "1"   # variable: A/B
        # larger batches: statistics are finalised by diqt_gn_finalize into (aff_a, aff_b) and the conv applies that affine + Mish
        self.fuse_gn_affine = (not self.grouped and self.sub_f <= 1 and self.dtype == "bf16" and os.environ.get("DIQT_DISABLE_GN_FUSION", "0") != "1")
        self.aff_a = torch.zeros(n * cmax, dtype=torch.float32, device=self.device)
        self.aff_b = torch.zeros(n * cmax, dtype=torch.float32, device=self.device)
        self.gate = torch.zeros(n * cmax, dtype=torch.float32, device=self.device)
        self._pp = 0  # ping-pong index of block-output partial buffers (part[0], part[1]); part[2] is intra-block scratch

        # ---- time conditioning weights
        th = u.to_time_hiddens
        self.w_four = self._f32(th[0].weights)
        self.w_t1, self.b_t1 = self._f32(th[1].weight), self._f32(th[1].bias)
        self.w_t2, self.b_t2 = self._f32(u.to_time_cond[0].weight), self._f32(u.to_time_cond[0].bias)
        self.tdim = u.time_cond_dim
        self._film_w, self._film_b, self._film_cols = [], [], 0

        def film_slot(block):
            off = self._film_cols
            self._film_w.append(block.time_mlp[1].weight.detach())
            self._film_b.append(block.time_mlp[1].bias.detach())
            self._film_cols += 2 * block.dim_out
            return off

        # ---- activation buffers per resolution level
        def act_buf(level, c):
            return Act(self._empty(n * self.level_vox[level], c), c, c)

        # concat buffers for the up path: Cat[l] lives at resolution level l and holds [up output (dims[l+1]) | skip (dims[l])]
        self.cat = {l: self._empty(n * self.level_vox[l], dims[l + 1] + dims[l]) for l in range(nl - 1)}

        # A (normalised input), H (conv output), R (res_conv output): shared by all blocks of a level
        scratch = {l: dict(c=need[l], A=self._empty(n * self.level_vox[l], need[l]), H=self._empty(n * self.level_vox[l], need[l]),
                           R=self._empty(n * self.level_vox[l], need[l])) for l in range(nl)}

        def level_scratch(level, c_need):
            assert scratch[level]["c"] >= c_need, (level, c_need, scratch[level]["c"])
            return scratch[level]

        pingpong = {}

        def next_out(level, c):
            key = (level, c)
            if key not in pingpong:
                pingpong[key] = [act_buf(level, c), act_buf(level, c), 0]
                self._keep += [pingpong[key][0].buf, pingpong[key][1].buf]
            pp = pingpong[key]
            pp[2] ^= 1
            return pp[pp[2]]

        gnf, aff, stats_fn, seg, sres = lib.diqt_gn_finalize, lib.diqt_affine_mish, lib.diqt_channel_stats, lib.diqt_se_gate, lib.diqt_scale_residual
        eng = self

        stats_g = lib.diqt_channel_stats_g

        def add_stats(x: Act, level, si):
            """A separate statistics pass over x into scratch set `si`; returns the stats record (partial, nblk, group, ngroups)."""
            nb, vox = self.nblk[level], self.level_vox[level]
            part = self.part[si]
            xp, xc, xl, pp = x.ptr, x.c, x.ld, part.data_ptr()
            sf, sh = self.level_sub[level]
            if self.grouped:
                ng = C.c_int(0)
                L.check(lib.diqt_stats_groups(nb, 1, C.byref(ng)), "stats_groups")
                gp_, tp_ = self.grp[si].data_ptr(), self.tick[si].data_ptr()
                ops.append(lambda st: L.check(stats_g(xp, dd, n, vox, xc, xl, nb, pp, gp_, tp_, st), "channel_stats_g"))
                return (part, nb, self.grp[si], ng.value)
            ops.append(lambda st: L.check(stats_fn(xp, dd, n, vox, xc, xl, nb, pp, sf, sh, st), "channel_stats"))
            return (part, nb, None, 0)

        def add_norm_act(x: Act, level, gn, film_off, dst_buf, name):
            """GroupNorm(+FiLM)+Mish of x into dst_buf[:, :x.c]; returns the Act."""
            part, nb, grp, ng = x.stats
            vox = self.level_vox[level]
            gamma, beta = self._f32(gn.weight), self._f32(gn.bias)
            groups, eps = gn.num_groups, float(gn.eps)
            gp, bp, c = gamma.data_ptr(), beta.data_ptr(), x.c
            dst = Act(dst_buf, x.c, x.c) if dst_buf is not None else None      # None: finalize only (the consumer applies the affine itself)
            xp, xl, nbk = x.ptr, x.ld, self.nblk_stream[level]
            dp, dl = (dst.ptr, dst.ld) if dst is not None else (0, 0)
            if grp is not None:
                # one kernel: GroupNorm finalisation in the prologue, then the affine + Mish pass
                gptr = grp.data_ptr()

                def op(st, film_off=film_off):
                    fptr = eng.film.data_ptr() + film_off * 4 if film_off is not None else 0
                    L.check(lib.diqt_gn_mish_g(xp, xl, dp, dl, dd, n, vox, c, gptr, ng, groups, eps, gp, bp, fptr, eng._film_cols if film_off is not None else 0,
                                               eng.film_row_ptr if film_off is not None else 0, eng.film_stride_n, nbk, st), name + ".gn_mish")
                ops.append(op)
                return dst
            pa, pb, pp = self.aff_a.data_ptr(), self.aff_b.data_ptr(), part.data_ptr()
            if film_off is None:
                ops.append(lambda st: L.check(gnf(pp, n, nb, vox, c, groups, eps, gp, bp, 0, 0, 0, 0, pa, pb, st), name + ".gn"))
            else:
                def op(st, film_off=film_off):
                    fptr = eng.film.data_ptr() + film_off * 4
                    L.check(gnf(pp, n, nb, vox, c, groups, eps, gp, bp, fptr, eng._film_cols, eng.film_row_ptr, eng.film_stride_n, pa, pb, st), name + ".gn")
                ops.append(op)
            if dst is None:
                return None
            sf, sh = self.level_sub[level]
            ops.append(lambda st: L.check(aff(xp, xl, dp, dl, dd, n, vox, c, pa, pb, nbk, sf, sh, st), name + ".mish"))
            return dst

        def add_resblock(blk, x: Act, level, out: Act, name):
            cin, cout = blk.dim, blk.dim_out
            assert x.c == cin and out.c == cout, (name, x.c, cin, out.c, cout)
            sc = level_scratch(level, max(cin, cout))
            vox = self.level_vox[level]
            if x.stats is None:
                x.stats = add_stats(x, level, self._pp)
                self._pp ^= 1
            film_off = film_slot(blk)

            def worth_fusing(ci):
                # measured (profiles/conv_gn_r2c.jsonl, r2q A/B): the fused conv beats apply + conv clearly from 32^3 x 128 channels up
                # (64^3 x 64: 51 vs 62 us), is level at 32^3 x 64 (whole step 2.214 vs 2.221 ms with those fused too) and loses on smaller
                # tensors, where the normalisation latency in front of the first MMA of every short z-segment exceeds the 5-8 us apply kernel
                return self.n * vox * ci >= self.gn_fusion_min

            def norm_conv(cname, src: Act, gmod, f_off, ci, dst: Act, weight, bias, si):
                """mish(FiLM(GroupNorm(src))) -> 3x3x3 conv -> dst; the normalisation rides on the conv's load path when the plan can."""
                desc = self._conv_desc(L.CONV_K3, level, ci, src.ld, cout, dst.ld)
                if not worth_fusing(ci):
                    pass
                elif self.fuse_gn and src.stats[2] is not None and lib.diqt_conv_gn_fusable(C.byref(desc)):
                    ops.append(self._conv_site(cname, L.CONV_K3, level, ci, src.ld, cout, dst.ld, weight, bias, src.ptr, dst.ptr, stats=si,
                                               gn=(src.stats, gmod, f_off)))
                    return True
                elif self.fuse_gn_affine and src.stats[2] is None and self._resolves_to_zm(desc):
                    add_norm_act(src, level, gmod, f_off, None, cname.rsplit(".", 1)[0])       # finalize only: (a, b) -> self.aff_a / aff_b
                    ops.append(self._conv_site(cname, L.CONV_K3, level, ci, src.ld, cout, dst.ld, weight, bias, src.ptr, dst.ptr, stats=si, gn="affine"))
                    return True
                a = add_norm_act(src, level, gmod, f_off, sc["A"], cname.rsplit(".", 1)[0])
                ops.append(self._conv_site(cname, L.CONV_K3, level, ci, a.ld, cout, dst.ld, weight, bias, a.ptr, dst.ptr, stats=si))
                return False

            h = Act(sc["H"], cout, cout)
            norm_conv(name + ".block1.project", x, blk.block1.groupnorm, None, cin, h, blk.block1.project.weight, blk.block1.project.bias, 2)
            h.stats = self._last_conv_stats or add_stats(h, level, 2)
            # conv2 cannot run in place: with the fused normalisation it reads h itself, so its output goes to the (otherwise unused) A buffer
            d2_ = self._conv_desc(L.CONV_K3, level, cout, h.ld, cout, cout)
            fusable2 = worth_fusing(cout) and ((self.fuse_gn and h.stats[2] is not None and lib.diqt_conv_gn_fusable(C.byref(d2_)))
                                               or (self.fuse_gn_affine and h.stats[2] is None and self._resolves_to_zm(d2_)))
            h2 = Act(sc["A"], cout, cout) if fusable2 else h
            norm_conv(name + ".block2.project", h, blk.block2.groupnorm, film_off, cout, h2, blk.block2.project.weight, blk.block2.project.bias,
                      3 if blk.has_se else None)
            h = h2
            gate_ptr, se = 0, None
            if blk.has_se:
                part, nb, grp, ng = self._last_conv_stats or add_stats(h, level, 3)
                w1, w2 = self._f32(blk.se.fc[0].weight), self._f32(blk.se.fc[2].weight)
                hidden = w1.shape[0]
                if hidden < 1:
                    raise ValueError(f"{name}: SE3D with {cout} channels has an empty bottleneck (reduction 16)")
                w1p, w2p = w1.data_ptr(), w2.data_ptr()
                if grp is not None:
                    se = (grp.data_ptr(), ng, hidden, w1p, w2p)       # gate computed in the residual kernel's prologue
                else:
                    gate_ptr = self.gate.data_ptr()
                    pp = part.data_ptr()
                    ops.append(lambda st: L.check(seg(pp, n, nb, vox, cout, hidden, w1p, w2p, gate_ptr, st), name + ".se"))
            if blk.has_res_conv:
                r = Act(sc["R"], cout, cout)
                ops.append(self._conv_site(name + ".res_conv", L.CONV_K1, level, cin, x.ld, cout, r.ld, blk.res_conv.weight, blk.res_conv.bias,
                                           x.ptr, r.ptr))
            else:
                r = x
            si = self._pp
            self._pp ^= 1
            opart = self.part[si]
            nbk = self.nblk[level]
            hp, hl, rp, rl, op_, ol, opp = h.ptr, h.ld, r.ptr, r.ld, out.ptr, out.ld, opart.data_ptr()
            sf, sh = self.level_sub[level]
            if self.grouped:
                ng_out = C.c_int(0)
                L.check(lib.diqt_stats_groups(nbk, 1, C.byref(ng_out)), "stats_groups")
                seg_, sng, shid, sw1, sw2 = se if se is not None else (0, 0, 0, 0, 0)
                if se is None and gate_ptr:
                    raise RuntimeError(f"{name}: grouped statistics expect the SE gate to come from grouped conv statistics")
                gop, tkp = self.grp[si].data_ptr(), self.tick[si].data_ptr()
                ops.append(lambda st: L.check(lib.diqt_scale_residual_g(hp, hl, rp, rl, op_, ol, dd, n, vox, cout, seg_, sng, shid, sw1, sw2, nbk, opp,
                                                                        gop, tkp, st), name + ".residual"))
                out.stats = (opart, nbk, self.grp[si], ng_out.value)
            else:
                ops.append(lambda st: L.check(sres(hp, hl, rp, rl, op_, ol, dd, n, vox, cout, gate_ptr, nbk, opp, sf, sh, st), name + ".residual"))
                out.stats = (opart, nbk, None, 0)
            if name in self.tap_names:
                self._add_tap(name, out, level)
            return out

        # ---- init conv
        x = next_out(0, dims[0])
        d0, d1, d2 = self.conv_dims[0]
        xp, xl, c0, nin, cn = x.ptr, x.ld, dims[0], u.init_channels, self.conv_n
        planes_ref, strides_ref = self._planes, self._pstrides
        sf0, sh0 = self.level_sub[0]
        cross_embed = hasattr(u.init_conv, "convs")
        self.init_conv_tc = (not cross_embed and self.dtype == "bf16" and self.sub_f <= 1 and 27 * nin <= 64 and c0 % 64 == 0 and self.impl != L.IMPL_SIMT
                             and os.environ.get("DIQT_DISABLE_INIT_TC", "0") != "1")
        self.init_conv_fused = (self.init_conv_tc and os.environ.get("DIQT_DISABLE_INIT_FUSED", "0") != "1"
                                and bool(lib.diqt_init_conv_tc_supported(nin, c0, d1, d2)))
        if self.init_conv_fused:
            # K = 27 * c_in = 54 is not a tensor-core shape as a 3x3x3 conv, but it is as a GEMM over im2col rows of K = 64: ONE persistent
            # kernel builds the rows in shared memory, multiplies them on the tensor cores and emits the first GroupNorm's statistics
            w = u.init_conv.weight.detach().to(self.device, torch.float32)                   # (c0, nin, 3, 3, 3)
            w2 = torch.zeros(c0, 64, device=self.device, dtype=torch.float32)
            w2[:, :27 * nin] = w.permute(0, 2, 3, 4, 1).reshape(c0, 27 * nin)                   # k = tap * nin + ci
            rows = torch.arange(c0, device=self.device)[:, None]
            chunks = torch.arange(8, device=self.device)[None, :]
            src = (chunks ^ (rows & 7))                                                      # position c holds logical chunk c ^ (row & 7)
            wsw = torch.gather(w2.to(torch.bfloat16).view(c0, 8, 8), 1, src[:, :, None].expand(c0, 8, 8)).contiguous()
            self._keep.append(wsw)
            bi = self._f32(u.init_conv.bias)
            nb_i = C.c_int(0)
            L.check(lib.diqt_init_conv_tc_blocks(cn, d0, d1, C.byref(nb_i)), "init_conv_tc_blocks")
            si = self._pp
            self._pp ^= 1
            part = self.part[si]
            assert part.numel() >= n * nb_i.value * c0 * 2
            ng_i = C.c_int(0)
            if self.grouped:
                L.check(lib.diqt_stats_groups(nb_i.value, 1, C.byref(ng_i)), "stats_groups")
            gp_, tp_ = (self.grp[si].data_ptr(), self.tick[si].data_ptr()) if self.grouped else (0, 0)
            wp, bp, pp = wsw.data_ptr(), bi.data_ptr(), part.data_ptr()
            ops.append(lambda st: L.check(lib.diqt_init_conv_tc(planes_ref, strides_ref, nin, wp, bp, xp, xl, cn, d0, d1, d2, c0, pp, gp_, tp_, st), "init_conv_tc"))
            x.stats = (part, nb_i.value, self.grp[si] if self.grouped else None, ng_i.value)
            self.conv_impls["init_conv"] = L.IMPL_TC
        elif self.init_conv_tc:
            # round-1 path (kept for shapes the fused kernel does not take): im2col rows in global memory + the 1x1x1 tcgen05 conv
            self.im2col = self._empty(n * vox0, 64, dtype=torch.bfloat16)
            w = u.init_conv.weight.detach().to(self.device, torch.float32)                   # (c0, nin, 3, 3, 3)
            w2 = torch.zeros(c0, 64, device=self.device, dtype=torch.float32)
            w2[:, :27 * nin] = w.permute(0, 2, 3, 4, 1).reshape(c0, 27 * nin)                   # k = tap * nin + ci
            colp = self.im2col.data_ptr()
            ops.append(lambda st: L.check(lib.diqt_init_im2col(planes_ref, strides_ref, nin, colp, cn, d0, d1, d2, st), "init_im2col"))
            ops.append(self._conv_site("init_conv", L.CONV_K1, 0, 64, 64, c0, xl, w2.reshape(c0, 64, 1, 1, 1), u.init_conv.bias, colp, xp,
                                       stats=self._pp))
            if self._last_conv_stats:
                x.stats = self._last_conv_stats
                self._pp ^= 1
        elif cross_embed:
            # CrossEmbedLayer (:661-686): one direct conv per kernel size into its channel slice of x
            co_off = 0
            for conv in u.init_conv.convs:
                k, nco = int(conv.kernel_size[0]), int(conv.out_channels)
                w = conv.weight.detach().to(self.device, torch.float32)                       # (nco, nin, k, k, k)
                wk = self._f32(w.permute(2, 3, 4, 1, 0).reshape(k ** 3, nin, nco).contiguous())   # [tap][ci][co]
                bk = self._f32(conv.bias)
                wp, bp = wk.data_ptr(), bk.data_ptr()
                ops.append(lambda st, k=k, nco=nco, wp=wp, bp=bp, co_off=co_off: L.check(
                    lib.diqt_init_conv_k(planes_ref, strides_ref, nin, k, wp, bp, xp, xl, co_off, nco, dd, cn, d0, d1, d2, st), f"init_conv.convs(k={k})"))
                co_off += nco
            assert co_off == c0, (co_off, c0)
        else:
            self.w_init = torch.empty(27 * u.init_channels * dims[0], dtype=torch.float32, device=self.device)
            wi = self._f32(u.init_conv.weight)
            self.b_init = self._f32(u.init_conv.bias)
            L.check(lib.diqt_init_conv_pack(wi.data_ptr(), dims[0], u.init_channels, self.w_init.data_ptr(), L.current_stream()), "init_conv_pack")
            wip, bip = self.w_init.data_ptr(), self.b_init.data_ptr()
            ops.append(lambda st: L.check(lib.diqt_init_conv(planes_ref, strides_ref, nin, wip, bip, xp, xl, dd, cn, d0, d1, d2, c0, sf0, sh0, st), "init_conv"))

        # ---- down path
        skip_scale = u.skip_connect_scale
        for l in range(nl):
            c = dims[l]
            level_blocks = [(u.downs[l][1], f"downs.{l}.1")] + [(b, f"downs.{l}.3.{i}") for i, b in enumerate(u.downs[l][3])]
            for j, (blk, name) in enumerate(level_blocks):
                last_of_level = j == len(level_blocks) - 1

                def level_out():
                    if last_of_level and l != nl - 1 and skip_scale == 1.0:
                        return Act(self.cat[l], c, dims[l + 1] + c, offset=dims[l + 1])  # write the skip straight into the concat buffer
                    return next_out(l, c)

                if j == 0 and u.downs[l][2] is not None:      # :1610-1622: x = attn(merge(x)) split again, x += res
                    x = add_resblock(blk, x, l, next_out(l, c), name)
                    x = self._add_attention(u.downs[l][2], x, l, level_out(), f"downs.{l}.2", outer_residual=True)
                else:
                    x = add_resblock(blk, x, l, level_out(), name)
            if l != nl - 1:
                if skip_scale != 1.0:
                    dst = Act(self.cat[l], c, dims[l + 1] + c, offset=dims[l + 1])
                    rows = n * self.level_vox[l]   # pitched row copy: layout (merged or not) does not matter
                    sp_, sl, dp_, dl = x.ptr, x.ld, dst.ptr, dst.ld
                    ops.append(lambda st, sp_=sp_, sl=sl, dp_=dp_, dl=dl, rows=rows, c=c: L.check(
                        lib.diqt_scale_copy(sp_, sl, dp_, dl, dd, rows, c, skip_scale, st), "skip_scale"))
                nxt = next_out(l + 1, dims[l + 1])
                conv = u.downs[l][4][1]
                ops.append(self._conv_site(f"downs.{l}.4.1", L.CONV_DOWN, l, c, x.ld, dims[l + 1], nxt.ld, conv.weight, conv.bias, x.ptr, nxt.ptr,
                                           stats=self._pp))
                if self._last_conv_stats:
                    nxt.stats = self._last_conv_stats
                    self._pp ^= 1
                x = nxt
            else:
                conv = u.downs[l][4]
                nxt = next_out(l, dims[l + 1])
                ops.append(self._conv_site(f"downs.{l}.4", L.CONV_K1, l, c, x.ld, dims[l + 1], nxt.ld, conv.weight, conv.bias, x.ptr, nxt.ptr,
                                           stats=self._pp))
                if self._last_conv_stats:
                    nxt.stats = self._last_conv_stats
                    self._pp ^= 1
                x = nxt

        level = nl - 1
        if u.deep_feature:
            if u.mid_attn is not None:                        # :1635-1646: no residual around mid_attn
                x = self._add_attention(u.mid_attn, x, level, next_out(level, dims[-1]), "mid_attn", outer_residual=False)
            x = add_resblock(u.mid_block, x, level, next_out(level, dims[-1]), "mid_block")

        # ---- up path
        for ui in range(nl):
            upsample, init_block, blocks = u.ups[ui]
            last = ui == nl - 1
            if not last:
                lo = nl - 2 - ui                      # resolution level of the output
                c_up = dims[lo + 1]
                ctot = c_up + dims[lo]
                cat = Act(self.cat[lo], ctot, ctot)
                if hasattr(upsample, "deconv"):
                    # Upsample_deconv (:440-457): ConvTranspose3d(k 3, stride 2) + Mish = a 3x3x3 conv to 8 c_up parity channels (zero taps
                    # where a parity does not reach), then Mish + pixel shuffle, which the pixel-shuffle conv applies with identity weights
                    wc, bc = deconv_as_conv3_weights(upsample.deconv[0].weight.detach().to(self.device), upsample.deconv[0].bias.detach().to(self.device))
                    par = self._empty(n * self.level_vox[lo + 1], 8 * c_up)
                    self._keep.append(par)
                    ops.append(self._conv_site(f"ups.{ui}.0.deconv.0", L.CONV_K3, lo + 1, x.c, x.ld, 8 * c_up, 8 * c_up, wc, bc, x.ptr, par.data_ptr()))
                    eye = torch.eye(8 * c_up, device=self.device).reshape(8 * c_up, 8 * c_up, 1, 1, 1)
                    ops.append(self._conv_site(f"ups.{ui}.0.deconv.shuffle", L.CONV_UP, lo + 1, 8 * c_up, 8 * c_up, 8 * c_up, ctot, eye, None, par.data_ptr(),
                                               cat.ptr))
                else:
                    conv = upsample.net[0]
                    # GEMM N = 8*c_up; stored channels c_up at pitch ctot (channel offset 0 of the concat buffer)
                    ops.append(self._conv_site(f"ups.{ui}.0.net.0", L.CONV_UP, lo + 1, x.c, x.ld, 8 * c_up, ctot, conv.weight, conv.bias, x.ptr, cat.ptr))
                x, level = cat, lo
            x = add_resblock(init_block, x, level, next_out(level, init_block.dim_out), f"ups.{ui}.1")
            for i, blk in enumerate(blocks):
                x = add_resblock(blk, x, level, next_out(level, blk.dim_out), f"ups.{ui}.2.{i}")

        if u.final_res_block is not None:
            x = add_resblock(u.final_res_block, x, 0, next_out(0, u.final_res_block.dim_out), "final_res_block")
        self.last_act = x
        self._scratch = scratch

        # ---- final conv
        self.w_final = self._f32(u.final_conv.weight.reshape(u.channels_out, -1))
        self.b_final = self._f32(u.final_conv.bias)
        self.pred = self._empty(n, u.channels_out, *sp, dtype=torch.float32)
        self.film_w = self._f32(torch.cat([w.to(self.device) for w in self._film_w], dim=0))
        self.film_b = self._f32(torch.cat([b.to(self.device) for b in self._film_b], dim=0))
        self.film = None
        torch.cuda.current_stream().synchronize()

    def _add_tap(self, name, act: Act, level: int):
        rows = self.n * self.level_vox[level]
        src = torch.as_strided(act.buf.view(-1), (rows, act.c), (act.ld, 1), act.off)
        dst = self._empty(rows, act.c)
        self.taps[name] = (dst, level)
        self._ops.append(lambda st: dst.copy_(src))

    def tap(self, name) -> torch.Tensor:
        """Tapped activation as the reference's (B, C, D, H, W) fp32 tensor (plain layout only)."""
        assert self.sub_f <= 1, "taps are recorded in the plain (non-boundary) layout"
        t, level = self.taps[name]
        d = self.level_dims[level]
        return t.view(self.n, *d, t.shape[1]).permute(0, 4, 1, 2, 3).float().contiguous()

    # ------------------------------------------------------------------ attention blocks
    def _add_attention(self, mod, x: Act, level: int, out: Act, name: str, outer_residual: bool) -> Act:
        """Append the launches of one attention block (imagen_pytorch3D.py:1610-1622 / 1635-1646) reading `x`, writing `out`.

        The reference merges the f^3 sub-volumes (utils_mine.py:44-67), runs the block on the merged volume and splits again.
        Here only the two kernels at the edges translate rows (`xmap`); in boundary mode the activations already live merged.
        1x1x1 convolutions run through the conv families (tcgen05 when the channel counts are multiples of 64), everything
        else through csrc/attn.cu."""
        lib, ops, dd, n, u = self.lib, self._ops, self.ddtype, self.n, self.unet
        f = int(u.batch_sample_factor)
        h = self.level_dims[level][0]
        if n != f ** 3 or len(set(self.level_dims[level])) != 1:
            # utils_mine.py:57-59 "The batch size must be the product of split dimensions"
            raise ValueError(f"{name}: attention needs batch == batch_sample_factor^3 = {f ** 3} cubic sub-volumes, got batch {n}, dims {self.level_dims[level]}")
        G, p, cdim = f * h, int(mod.patch_size), int(mod.dim)
        if h % p != 0:
            raise ValueError(f"{name}: sub-volume side {h} is not a multiple of the attention patch size {p}")
        g = G // p
        N = g ** 3
        rows = n * self.level_vox[level]
        assert x.c == cdim and out.c == cdim, (name, x.c, out.c, cdim)
        xmap = (0, 0) if self.sub_f > 1 or f == 1 else (f, h)
        heads, dh = int(mod.heads), int(mod.dim_head)
        inner = heads * dh
        impl = L.IMPL_SIMT if self.dtype == "fp32" else L.IMPL_AUTO
        tok_vol = (1, g, g, g)

        def buf(r, c):
            t = self._empty(r, c)
            self._keep.append(t)
            return Act(t, c, c)

        def dw_weight(conv):
            w = conv.weight.detach().to(device=self.device, dtype=torch.float32)
            return self._f32(w.reshape(w.shape[0], -1).t().contiguous())

        def opt_f32(t):
            return self._f32(t) if t is not None else None

        def ln(src: Act, dst: Act, r, gain, beta=None, eps=1e-5, pre_act=0, res1=None, res2=None, merged_src=False, what="ln"):
            gp, bp = self._f32(gain.reshape(-1)).data_ptr(), L.ptr(opt_f32(beta))
            sp, sl, dp, dl, c = src.ptr, src.ld, dst.ptr, dst.ld, src.c
            r1p, r1l = (res1.ptr, res1.ld) if res1 is not None else (0, 0)
            r2p, r2l = (res2.ptr, res2.ld) if res2 is not None else (0, 0)
            sf, sh = xmap if merged_src else (0, 0)
            ops.append(lambda st: L.check(lib.diqt_chan_layernorm(sp, sl, dp, dl, dd, r, c, gp, bp, eps, pre_act, r1p, r1l, r2p, r2l, sf, sh, st),
                                          f"{name}.{what}"))

        def combine(a: Act, dst: Act, r, act=0, b=None, c2=None, what="add"):
            ap, al, dp, dl, c = a.ptr, a.ld, dst.ptr, dst.ld, a.c
            bp, bl = (b.ptr, b.ld) if b is not None else (0, 0)
            cp, cl = (c2.ptr, c2.ld) if c2 is not None else (0, 0)
            ops.append(lambda st: L.check(lib.diqt_rows_combine(ap, al, act, bp, bl, cp, cl, dp, dl, dd, r, c, st), f"{name}.{what}"))

        def conv1(src: Act, dst: Act, weight, bias, what, vol=None, c_out=None):
            co = c_out if c_out is not None else dst.c
            w = weight.detach().reshape(co, src.c, 1, 1, 1)
            ops.append(self._conv_site(f"{name}.{what}", L.CONV_K1, level, src.c, src.ld, co, dst.ld, w, bias, src.ptr, dst.ptr, vol=vol, impl=impl))

        def dwconv3(src: Act, dst: Act, conv, side, what):
            wp, bp = dw_weight(conv).data_ptr(), L.ptr(opt_f32(conv.bias))
            sp, sl, dp, dl, c = src.ptr, src.ld, dst.ptr, dst.ld, src.c
            ops.append(lambda st: L.check(lib.diqt_dw_conv3(sp, sl, dp, dl, dd, side, side, side, c, wp, bp, st), f"{name}.{what}"))

        def patchify(src: Act, dst: Act, conv, what):
            wp, bp = dw_weight(conv).data_ptr(), L.ptr(opt_f32(conv.bias))
            sp, sl, dp, dl = src.ptr, src.ld, dst.ptr, dst.ld
            ops.append(lambda st: L.check(lib.diqt_dw_patchify(sp, sl, dp, dl, dd, g, p, cdim, wp, bp, xmap[0], xmap[1], st), f"{name}.{what}"))

        def upsample(src: Act, dst: Act, what):
            sp, sl, dp, dl = src.ptr, src.ld, dst.ptr, dst.ld
            ops.append(lambda st: L.check(lib.diqt_upsample_trilinear(sp, sl, dp, dl, dd, g, p, cdim, st), f"{name}.{what}"))

        def softmax_attn(dst: Act, act: int, what: str):
            """softmax_k(q k^T * scale) v per head into dst: the tcgen05 kernel when the shape allows (bf16, dim_head 64), else CUDA cores."""
            dp, dl = dst.ptr, dst.ld
            if (self.dtype == "bf16" and os.environ.get("DIQT_DISABLE_ATTN_TC", "0") != "1"
                    and lib.diqt_attn_tc_supported(dd, dh, ldq, ldq, ldq, dl)):
                nbytes = C.c_size_t(0)
                L.check(lib.diqt_attn_tc_workspace_bytes(N, heads, C.byref(nbytes)), "attn_tc_workspace_bytes")
                ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.device)      # V^T, padding columns stay zero
                self._keep.append(ws)
                plan = C.c_void_p(0)
                L.check(lib.diqt_attn_tc_plan_create(q_ptr, k_ptr, v_ptr, ldq, ldq, ldq, dp, dl, N, heads, scale, act, ws.data_ptr(), C.byref(plan)),
                        f"{name}.{what}.plan")
                self._attn_plans.append(plan.value)
                pv = plan.value
                self.attn_impls[f"{name}.{what}"] = "tc"
                ops.append(lambda st: L.check(lib.diqt_attn_tc_run(pv, st), f"{name}.{what}"))
            else:
                self.attn_impls[f"{name}.{what}"] = "simt"
                ops.append(lambda st: L.check(lib.diqt_softmax_attention(q_ptr, k_ptr, v_ptr, ldq, ldq, ldq, dp, dl, dd, N, heads, dh, scale, act, st),
                                              f"{name}.{what}"))

        esz = 2 if self.dtype == "bf16" else 4
        F0, F1, F2 = buf(rows, cdim), buf(rows, cdim), buf(rows, cdim)
        T0, T1, T2, T3 = buf(N, cdim), buf(N, cdim), buf(N, cdim), buf(N, cdim)
        QKV0, QKV, O = buf(N, 3 * inner), buf(N, 3 * inner), buf(N, inner)
        q_ptr, k_ptr, v_ptr, ldq = QKV.ptr, QKV.ptr + inner * esz, QKV.ptr + 2 * inner * esz, 3 * inner
        scale = float(dh) ** -0.5

        if mod.kind == "vit":
            # ---- ViT3D.forward :905-910
            pe = mod.patch_embedding
            patchify(x, T0, pe.projection[0].depthwise, "patch_embedding.depthwise")
            conv1(T0, T1, pe.projection[0].pointwise.weight, pe.projection[0].pointwise.bias, "patch_embedding.pointwise", vol=tok_vol)
            if tuple(pe.positions.shape) != (N, cdim):
                raise ValueError(f"{name}: position table {tuple(pe.positions.shape)} does not match {N} tokens x {cdim} (img_size / patch_size of the constructor)")
            pos_t = pe.positions.detach().to(device=self.device, dtype=self.tdtype).contiguous()
            self._keep.append(pos_t)
            tok = buf(N, cdim)
            combine(T1, tok, N, b=Act(pos_t, cdim, cdim), what="positions")
            hid = cdim * int(mod.expansion)
            U0, U1 = buf(N, hid), buf(N, hid)
            for i, layer in enumerate(mod.transformer_encoder.layers):
                ln_a, mha = layer.block[0].fn[0], layer.block[0].fn[1]
                ln(tok, T2, N, ln_a.weight, ln_a.bias, eps=float(ln_a.eps), what=f"layers.{i}.ln1")
                # 'b n (h d qkv) -> qkv b h n d' (:824): permute the rows of the projection so that q | k | v come out as blocks
                wq = mha.qkv.weight.detach().reshape(heads, dh, 3, cdim).permute(2, 0, 1, 3).reshape(3 * inner, cdim)
                bq = mha.qkv.bias.detach().reshape(heads, dh, 3).permute(2, 0, 1).reshape(3 * inner)
                conv1(T2, QKV, wq, bq, f"layers.{i}.qkv", vol=tok_vol, c_out=3 * inner)
                softmax_attn(O, 0, f"layers.{i}.mha")
                conv1(O, T3, mha.projection.weight, mha.projection.bias, f"layers.{i}.projection", vol=tok_vol)
                tok2 = buf(N, cdim)
                combine(T3, tok2, N, b=tok, what=f"layers.{i}.res1")
                ln_f, ffb = layer.block[1].fn[0], layer.block[1].fn[1]
                ln(tok2, T2, N, ln_f.weight, ln_f.bias, eps=float(ln_f.eps), what=f"layers.{i}.ln2")
                if ffb.local:
                    conv1(T2, U0, ffb.up_proj[1].weight, ffb.up_proj[1].bias, f"layers.{i}.up_proj", vol=tok_vol)
                    combine(U0, U1, N, act=1, what=f"layers.{i}.up_mish")
                    dwconv3(U1, U0, ffb.depth_conv[0].depthwise, g, f"layers.{i}.depth_conv.depthwise")
                    conv1(U0, U1, ffb.depth_conv[0].pointwise.weight, ffb.depth_conv[0].pointwise.bias, f"layers.{i}.depth_conv.pointwise", vol=tok_vol)
                    combine(U1, U0, N, act=1, what=f"layers.{i}.depth_mish")
                    conv1(U0, T3, ffb.down_proj[0].weight, ffb.down_proj[0].bias, f"layers.{i}.down_proj", vol=tok_vol)
                else:
                    conv1(T2, U0, ffb.net[0].weight, ffb.net[0].bias, f"layers.{i}.ff0", vol=tok_vol)
                    combine(U0, U1, N, act=1, what=f"layers.{i}.ff_mish")
                    conv1(U1, T3, ffb.net[3].weight, ffb.net[3].bias, f"layers.{i}.ff1", vol=tok_vol)
                tok = buf(N, cdim)
                combine(T3, tok, N, b=tok2, what=f"layers.{i}.res2")
            rec = mod.reconstruction
            ln(tok, T2, N, rec[0].weight, rec[0].bias, eps=float(rec[0].eps), what="reconstruction.ln")
            upsample(T2, F1, "reconstruction.upsample")
            dwconv3(F1, F0, rec[3].depthwise, G, "reconstruction.depthwise")
            conv1(F0, F1, rec[3].pointwise.weight, rec[3].pointwise.bias, "reconstruction.pointwise")
            ln(F1, out, rows, rec[4].g, res1=x if outer_residual else None, merged_src=True, what="reconstruction.norm")
            out.stats = None
            return out

        # ---- {Linear,SoftMax}AttentionTransformerBlock.forward :1146-1150 / :1181-1186
        lin_tc = (mod.kind == "linear" and self.dtype == "bf16" and os.environ.get("DIQT_DISABLE_ATTN_TC", "0") != "1"
                  and bool(lib.diqt_linattn_tc_supported(dd, dh, heads, 3 * inner, inner)))
        if mod.kind == "linear" and lin_tc:
            nbytes = C.c_size_t(0)
            L.check(lib.diqt_linattn_tc_workspace_bytes(N, heads, C.byref(nbytes)), "linattn_tc_workspace_bytes")
            lin_ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=self.device)
            self._keep.append(lin_ws)
        elif mod.kind == "linear":
            nch = C.c_int(0)
            L.check(lib.diqt_linear_attention_chunks(N, C.byref(nch)), "linear_attention_chunks")
            col_stat = torch.zeros(inner * 2, dtype=torch.float32, device=self.device)
            ctx_part = torch.zeros(nch.value * heads * dh * dh, dtype=torch.float32, device=self.device)
            self._keep += [col_stat, ctx_part]
        depth = len(mod.layers)
        hidden = mod.layers[0][1][1].weight.shape[0]
        H0, H1 = buf(rows, hidden), buf(rows, hidden)
        mid = buf(rows, cdim) if depth > 1 else None
        xin = x
        for i, (attn, ff) in enumerate(mod.layers):
            last = i == depth - 1
            lname = f"layers.{i}"
            # Patchify :926-929 -> norm -> q, k, v (1x1x1 then depthwise 3x3x3, no biases) :961-977
            ln(xin, F0, rows, attn.patch_embed.norm.g, what=f"{lname}.patch_embed.norm")
            patchify(F0, T0, attn.patch_embed.projection.depthwise, f"{lname}.patch_embed.depthwise")
            conv1(T0, T1, attn.patch_embed.projection.pointwise.weight, attn.patch_embed.projection.pointwise.bias, f"{lname}.patch_embed.pointwise", vol=tok_vol)
            ln(T1, T2, N, attn.norm.g, what=f"{lname}.norm")
            wqkv = torch.cat([attn.to_q[1].weight.detach(), attn.to_k[1].weight.detach(), attn.to_v[1].weight.detach()], dim=0)
            conv1(T2, QKV0, wqkv, None, f"{lname}.to_qkv.1", vol=tok_vol, c_out=3 * inner)
            wdw = torch.cat([attn.to_q[2].weight.detach(), attn.to_k[2].weight.detach(), attn.to_v[2].weight.detach()], dim=0)
            wdw = self._f32(wdw.to(self.device, torch.float32).reshape(3 * inner, 27).t().contiguous())
            wp, s0, s0l, d0p, d0l = wdw.data_ptr(), QKV0.ptr, QKV0.ld, QKV.ptr, QKV.ld
            ops.append(lambda st, wp=wp, lname=lname: L.check(lib.diqt_dw_conv3(s0, s0l, d0p, d0l, dd, g, g, g, 3 * inner, wp, 0, st), f"{name}.{lname}.to_qkv.2"))
            op_, ol = O.ptr, O.ld
            if mod.kind == "linear" and lin_tc:
                plan = C.c_void_p(0)     # the layers of a block run one after the other: they share the workspace
                L.check(lib.diqt_linattn_tc_plan_create(q_ptr, k_ptr, v_ptr, ldq, op_, ol, N, heads, scale, 1, (lin_ws.data_ptr() + 255) // 256 * 256,
                                                        C.byref(plan)), f"{name}.{lname}.linear_attention.plan")
                self._linattn_plans.append(plan.value)
                self.attn_impls[f"{name}.{lname}.linear_attention"] = "tc"
                ops.append(lambda st, pv=plan.value, lname=lname: L.check(lib.diqt_linattn_tc_run(pv, st), f"{name}.{lname}.linear_attention"))
            elif mod.kind == "linear":
                self.attn_impls[f"{name}.{lname}.linear_attention"] = "simt"
                csp, cpp = col_stat.data_ptr(), ctx_part.data_ptr()
                ops.append(lambda st, lname=lname: L.check(lib.diqt_linear_attention(q_ptr, k_ptr, v_ptr, ldq, op_, ol, dd, N, heads, dh, scale, 1, csp, cpp, st),
                                              f"{name}.{lname}.linear_attention"))
            else:
                softmax_attn(O, 1, f"{lname}.softmax_attention")
            conv1(O, T3, attn.to_out[0].weight, None, f"{lname}.to_out.0", vol=tok_vol)
            ln(T3, T0, N, attn.to_out[1].g, what=f"{lname}.to_out.1")
            # reconstruct :952-959, then "+ x" :1148
            upsample(T0, F1, f"{lname}.reconstruct.upsample")
            dwconv3(F1, F0, attn.reconstruct[1].depthwise, G, f"{lname}.reconstruct.depthwise")
            conv1(F0, F1, attn.reconstruct[1].pointwise.weight, attn.reconstruct[1].pointwise.bias, f"{lname}.reconstruct.pointwise")
            ln(F1, F2, rows, attn.reconstruct[2].g, res1=xin, merged_src=True, what=f"{lname}.reconstruct.norm")
            # ChanFeedForward :1108-1116, then "+ x" :1149 (and the U-Net's own "x += res" :1622 after the last layer)
            ln(F2, F0, rows, ff[0].g, what=f"{lname}.ff.0")
            conv1(F0, H0, ff[1].weight, None, f"{lname}.ff.1")
            ln(H0, H1, rows, ff[3].g, pre_act=2, what=f"{lname}.ff.3")
            conv1(H1, F0, ff[4].weight, None, f"{lname}.ff.4")
            dst = out if last else mid
            combine(F0, dst, rows, b=F2, c2=x if (last and outer_residual) else None, what=f"{lname}.residual")
            xin = dst
        out.stats = None
        return out

    # ------------------------------------------------------------------ conditioning
    def set_condition(self, log_snr: torch.Tensor, stream=None):
        """Evaluate the time MLPs for `log_snr` (R,) -> FiLM table [R][sum 2C] (one row per batch element for
        a plain forward, one row per sampler step in the sampler)."""
        lib = self.lib
        st = stream if stream is not None else L.current_stream()
        rows = int(log_snr.shape[0])
        t = log_snr.detach().to(device=self.device, dtype=torch.float32).contiguous()
        half = self.w_four.shape[0]
        if self.film is None or self._film_rows < rows:
            # captured step graphs bake self.film.data_ptr(): a larger table means new buffers, so the generation counter tells every
            # cached sampler state to drop its graph and capture again on next use (imagen.py / elucidated.py compare it)
            self.film_gen += 1
            self._film_rows = rows
            self.four = torch.empty(rows, 1 + 2 * half, dtype=torch.float32, device=self.device)
            self.thid = torch.empty(rows, self.tdim, dtype=torch.float32, device=self.device)
            self.tcond = torch.empty(rows, self.tdim, dtype=torch.float32, device=self.device)
            self.film = torch.empty(rows, self._film_cols, dtype=torch.float32, device=self.device)
        self._t_keep = t
        L.check(lib.diqt_fourier_features(t.data_ptr(), rows, self.w_four.data_ptr(), half, self.four.data_ptr(), st), "fourier")
        L.check(lib.diqt_linear(self.four.data_ptr(), 1 + 2 * half, rows, 1 + 2 * half, self.w_t1.data_ptr(), self.b_t1.data_ptr(), self.tdim,
                                self.thid.data_ptr(), self.tdim, 0, 1, st), "to_time_hiddens")
        L.check(lib.diqt_linear(self.thid.data_ptr(), self.tdim, rows, self.tdim, self.w_t2.data_ptr(), self.b_t2.data_ptr(), self.tdim,
                                self.tcond.data_ptr(), self.tdim, 0, 0, st), "to_time_cond")
        L.check(lib.diqt_linear(self.tcond.data_ptr(), self.tdim, rows, self.tdim, self.film_w.data_ptr(), self.film_b.data_ptr(), self._film_cols,
                                self.film.data_ptr(), self._film_cols, 1, 0, st), "time_mlps")

    # ------------------------------------------------------------------ execution
    def run_body(self, stream=None):
        st = stream if stream is not None else L.current_stream()
        for op in self._ops:
            op(st)

    def run_final(self, stream=None, *, fused=False, sched=0, step=0, noise=0, x0=0):
        """final 1x1x1 conv; fused=True also applies the DDPM update in place on x_in."""
        st = stream if stream is not None else L.current_stream()
        x, u = self.last_act, self.unet
        sf, sh = self.level_sub[0]
        L.check(self.lib.diqt_final_conv(x.ptr, x.ld, self.ddtype, self.conv_n, self.level_vox[0] * (self.n // self.conv_n), x.c, u.channels_out,
                                         self.w_final.data_ptr(), self.b_final.data_ptr(), self.pred.data_ptr(), 1 if fused else 0, sched, step,
                                         self.x_in.data_ptr(), noise, self.x_in.data_ptr(), x0, sf, sh, st), "final_conv")

    def load_inputs(self, x=None, lowres_cond_img=None, cond_images=None):
        if x is not None:
            self.x_in.copy_(x, non_blocking=True)
        if self.lowres is not None and lowres_cond_img is not None:
            self.lowres.copy_(lowres_cond_img, non_blocking=True)
        if self.cond_images is not None and cond_images is not None:
            self.cond_images.copy_(cond_images, non_blocking=True)

    def forward(self, x, time, *, lowres_cond_img=None, cond_images=None, self_cond=None):
        assert tuple(x.shape) == tuple(self.x_in.shape), (x.shape, self.x_in.shape)
        assert time.shape[0] == self.n
        self.load_inputs(x, lowres_cond_img, cond_images)
        self.set_condition(time)
        self.film_row_ptr, self.film_stride_n = 0, 1
        self.run_body()
        self.run_final()
        return self.pred.clone()

    def close(self):
        for p in self._plans:
            self.lib.diqt_conv_plan_destroy(p)
        self._plans = []
        for p in getattr(self, "_attn_plans", []):
            self.lib.diqt_attn_tc_plan_destroy(p)
        self._attn_plans = []
        for p in getattr(self, "_linattn_plans", []):
            self.lib.diqt_linattn_tc_plan_destroy(p)
        self._linattn_plans = []
        self._ops = []
        self.sampler_cache = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
