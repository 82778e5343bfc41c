"""BASELINE config 5: kernel sweep of this library against the PyTorch ops the reference calls, on the same GPU.

  conv   : Conv3d 3x3x3, C -> C, C in {64, 128, 256, 512} at {16^3, 32^3, 64^3}        F.conv3d (cuDNN) in bf16 channels_last_3d,
                                                                                        bf16 contiguous (NCDHW) and fp32 with TF32
  norm   : GroupNorm(8) + FiLM + Mish on the same grid                                  F.group_norm -> x*(s+1)+t -> F.mish, bf16
  attn   : softmax attention, 8 heads x 64, N in {1728, 13824} tokens                   F.scaled_dot_product_attention, bf16

Every timing: `reps` launches over rotating buffers (so consecutive launches do not hit the same L2 lines) captured as ONE CUDA graph
when the op can be captured (ours always; torch ops after a warm-up), CUDA events around 3 replays.  One JSON line per (op, shape);
`speedup` = best torch time / our time.  No oracle, no reference tree: only torch ops and libdiqt_b200.
Usage: python tools/bench_sweep.py [conv] [norm] [attn] > profiles/sweep_rN.jsonl"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

import bench
from diffusioniqt_b200 import lib as L

lib = L.load()
peaks = bench.read_peaks()
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True            # let cuDNN pick its best algorithm: the strongest baseline


def timed(fns, reps):
    """fns: list of callables (one per rotating buffer set).  Returns ms per call."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = None
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for f in fns:
                f()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                fns[i % len(fns)]()
    except Exception as e:                        # not capturable: time eager launches
        g = None
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if g is not None:
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (3 * reps), "graph"
    e0.record()
    for i in range(3 * reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps), "eager"


def nbuf_for(bytes_per_set):
    return max(2, min(6, int(400e6 // max(bytes_per_set, 1)) + 1))


def sweep_conv():
    for S in (16, 32, 64):
        for Cc in (64, 128, 256, 512):
            vox = S ** 3
            flops = 2.0 * Cc * Cc * 27 * vox
            nb = nbuf_for(2 * vox * Cc * 2)
            reps = 20 if flops < 5e11 else 6
            w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) * (27 * Cc) ** -0.5
            b = torch.randn(Cc, device=dev) * 0.1
            # ---- ours
            xs = [torch.randn(1, S, S, S, Cc, device=dev).bfloat16() for _ in range(nb)]
            ys = [torch.empty(1, S, S, S, Cc, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
            def ours(impl_id, flags=0, split_k=True):
                desc = L.ConvDesc(mode=L.CONV_K3, dtype=L.BF16, impl=impl_id, n=1, d0=S, d1=S, d2=S, c_in=Cc, ld_in=Cc, c_out=Cc, ld_out=Cc, flags=flags)
                impl = C.c_int(0)
                L.check(lib.diqt_conv_resolved_impl(C.byref(desc), C.byref(impl)))
                nbytes = C.c_size_t(0)
                L.check(lib.diqt_conv_packed_bytes(C.byref(desc), C.byref(nbytes)))
                packed = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
                pb = torch.empty(Cc, dtype=torch.float32, device=dev)
                L.check(lib.diqt_conv_pack(C.byref(desc), wb.data_ptr(), b.data_ptr(), packed.data_ptr(), pb.data_ptr(), L.current_stream()))
                plans, keep = [], []
                for x, y in zip(xs, ys):
                    p = C.c_void_p(0)
                    L.check(lib.diqt_conv_plan_create(C.byref(desc), x.data_ptr(), y.data_ptr(), packed.data_ptr(), pb.data_ptr(), C.byref(p)))
                    wsb = C.c_size_t(0)
                    L.check(lib.diqt_conv_plan_workspace_bytes(p.value, C.byref(wsb)))
                    if wsb.value and split_k:
                        ws = torch.zeros(wsb.value, dtype=torch.uint8, device=dev)
                        keep.append(ws)
                        L.check(lib.diqt_conv_plan_set_workspace(p.value, ws.data_ptr(), wsb.value))
                    plans.append(p.value)
                t, _ = timed([(lambda p=p: L.check(lib.diqt_conv_run(p, L.current_stream()))) for p in plans], reps)
                torch.cuda.synchronize()
                for p in plans:
                    lib.diqt_conv_plan_destroy(p)
                return t, impl.value

            wb = w.bfloat16().float().contiguous()
            alt = {}
            alt["conv_tc_kernel (per tap, no split-K)"], _ = ours(L.IMPL_TC, split_k=False)
            alt["conv_zm_kernel single CTA"], _ = ours(L.IMPL_ZM, L.CONV_FLAG_NO_CTA_PAIR)
            ours_ms, impl_v = ours(L.IMPL_AUTO)          # last: ys[0] holds the product path's output for the cross-check below
            impl = C.c_int(impl_v)
            ours = ours_ms
            # numerical cross-check against cuDNN on the first buffer (bf16 in, fp32 accumulate on both sides)
            ref = F.conv3d(xs[0].permute(0, 4, 1, 2, 3).float(), wb, b, padding=1)
            err = ((ys[0].permute(0, 4, 1, 2, 3).float() - ref).abs().max() / ref.abs().max()).item()
            del ref
            # ---- torch
            t = {}
            wcl = w.bfloat16().contiguous(memory_format=torch.channels_last_3d)
            xcl = [x.permute(0, 4, 1, 2, 3) for x in xs]                            # NDHWC storage viewed as NCDHW == channels_last_3d
            assert xcl[0].is_contiguous(memory_format=torch.channels_last_3d)
            bb = b.bfloat16()
            t["bf16_channels_last_3d"], how = timed([(lambda x=x: F.conv3d(x, wcl, bb, padding=1)) for x in xcl], reps)
            xnc = [x.permute(0, 4, 1, 2, 3).contiguous() for x in xs]
            wnc = w.bfloat16()
            t["bf16_ncdhw"], _ = timed([(lambda x=x: F.conv3d(x, wnc, bb, padding=1)) for x in xnc], reps)
            del xnc
            if vox * Cc * 4 * nb < 3e9:
                x32 = [x.permute(0, 4, 1, 2, 3).float().contiguous() for x in xs]
                torch.backends.cudnn.allow_tf32 = True
                t["fp32_tf32_ncdhw"], _ = timed([(lambda x=x: F.conv3d(x, w, b, padding=1)) for x in x32], reps)
                del x32
            best = min(t.values())
            kernel = {L.IMPL_SIMT: "conv_simt_kernel", L.IMPL_TC: "conv_tc_kernel", L.IMPL_ZM: "conv_zm_kernel"}[impl.value]
            print(json.dumps(dict(op="conv3x3x3", side=S, channels=Cc, gflop=flops / 1e9, ours_ms=ours, ours_kernel=kernel, ours_other_kernels_ms=alt,
                                  ours_tflops=flops / ours / 1e9, ours_frac_burst=flops / ours / 1e9 / peaks["burst"], torch_ms=t,
                                  torch_best_tflops=flops / best / 1e9, speedup=best / ours, max_rel_vs_cudnn=err, torch_timing=how)), flush=True)
            del xs, ys, xcl
            torch.cuda.empty_cache()


def sweep_norm():
    for S in (16, 32, 64):
        for Cc in (64, 128, 256, 512):
            vox = S ** 3
            nb = nbuf_for(2 * vox * Cc * 2)
            reps = 20
            xs = [torch.randn(vox, Cc, device=dev).bfloat16() for _ in range(nb)]
            ys = [torch.empty(vox, Cc, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
            gamma, beta = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
            film = (torch.randn(1, 2 * Cc, device=dev) * 0.2).contiguous()
            nblk_s, nblk_a = max(1, min(vox // 128, 148)), max(1, min(vox // 128, 592))
            ng = C.c_int(0)
            L.check(lib.diqt_stats_groups(nblk_s, 1, C.byref(ng)))
            part = torch.zeros(nblk_s * Cc * 2, device=dev)
            grp = torch.zeros(16 * Cc * 2, device=dev)
            tick = torch.zeros(16, dtype=torch.int32, device=dev)

            def stats(x):
                L.check(lib.diqt_channel_stats_g(x.data_ptr(), L.BF16, 1, vox, Cc, Cc, nblk_s, part.data_ptr(), grp.data_ptr(), tick.data_ptr(), L.current_stream()))

            def apply(x, y):
                L.check(lib.diqt_gn_mish_g(x.data_ptr(), Cc, y.data_ptr(), Cc, L.BF16, 1, vox, Cc, grp.data_ptr(), ng.value, 8, 1e-5, gamma.data_ptr(),
                                           beta.data_ptr(), film.data_ptr(), 2 * Cc, 0, 1, nblk_a, L.current_stream()))

            both, _ = timed([(lambda x=x, y=y: (stats(x), apply(x, y))) for x, y in zip(xs, ys)], reps)
            only_apply, _ = timed([(lambda x=x, y=y: apply(x, y)) for x, y in zip(xs, ys)], reps)
            # torch: bf16, channels_last_3d storage (what our buffers are) and contiguous NCDHW (what the reference holds)
            gb, bb = gamma.bfloat16(), beta.bfloat16()
            sc = (film[0, :Cc] + 1).bfloat16().view(1, Cc, 1, 1, 1)
            sh = film[0, Cc:].bfloat16().view(1, Cc, 1, 1, 1)

            def ref(x):
                return F.mish(F.group_norm(x, 8, gb, bb, 1e-5) * sc + sh)

            t = {}
            xcl = [x.view(1, S, S, S, Cc).permute(0, 4, 1, 2, 3) for x in xs]
            t["bf16_channels_last_3d"], how = timed([(lambda x=x: ref(x)) for x in xcl], reps)
            xnc = [x.contiguous() for x in xcl]
            t["bf16_ncdhw"], _ = timed([(lambda x=x: ref(x)) for x in xnc], reps)
            want = F.mish(F.group_norm(xnc[0].float(), 8, gamma, beta, 1e-5) * (film[0, :Cc] + 1).view(1, Cc, 1, 1, 1) + film[0, Cc:].view(1, Cc, 1, 1, 1))
            stats(xs[0]); apply(xs[0], ys[0]); torch.cuda.synchronize()
            err = ((ys[0].view(1, S, S, S, Cc).permute(0, 4, 1, 2, 3).float() - want).abs().max() / want.abs().max()).item()
            best = min(t.values())
            nbytes = 2.0 * vox * Cc * 2
            print(json.dumps(dict(op="groupnorm8+film+mish", side=S, channels=Cc, ours_stats_plus_apply_ms=both, ours_apply_only_ms=only_apply,
                                  ours_apply_gbs=nbytes / only_apply / 1e6, ours_apply_frac_hbm=nbytes / only_apply / 1e6 / peaks["hbm"], torch_ms=t,
                                  speedup_vs_stats_plus_apply=best / both, speedup_vs_apply_only=best / only_apply, max_rel_vs_torch_fp32=err,
                                  torch_timing=how)), flush=True)
            del xs, ys, xcl, xnc
            torch.cuda.empty_cache()


def sweep_attn():
    heads, dh = 8, 64
    inner = heads * dh
    for n in (1728, 13824):
        qkvs = [torch.randn(n, 3 * inner, device=dev).bfloat16() for _ in range(3)]
        outs = [torch.empty(n, inner, dtype=torch.bfloat16, device=dev) for _ in range(3)]
        nbytes = C.c_size_t(0)
        L.check(lib.diqt_attn_tc_workspace_bytes(n, heads, C.byref(nbytes)))
        plans, keep = [], []
        for qkv, out in zip(qkvs, outs):
            ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
            keep.append(ws)
            p, plan = qkv.data_ptr(), C.c_void_p(0)
            L.check(lib.diqt_attn_tc_plan_create(p, p + inner * 2, p + 2 * inner * 2, 3 * inner, 3 * inner, 3 * inner, out.data_ptr(), inner, n, heads,
                                                 dh ** -0.5, 0, ws.data_ptr(), C.byref(plan)))
            plans.append(plan.value)
        reps = 20 if n < 4000 else 6
        ours, _ = timed([(lambda p=p: L.check(lib.diqt_attn_tc_run(p, L.current_stream()))) for p in plans], reps)
        qs = [q.view(n, 3, heads, dh).permute(1, 2, 0, 3).unsqueeze(1).contiguous() for q in qkvs]       # (3, 1, heads, n, dh)
        t = {}
        t["sdpa_bf16"], how = timed([(lambda q=q: F.scaled_dot_product_attention(q[0], q[1], q[2])) for q in qs], reps)
        want = F.scaled_dot_product_attention(qs[0][0].float(), qs[0][1].float(), qs[0][2].float())[0]   # (heads, n, dh)
        got = outs[0].view(n, heads, dh).permute(1, 0, 2).float()
        err = ((got - want).abs().max() / want.abs().max()).item()
        fl = 4.0 * n * n * dh * heads
        for p in plans:
            lib.diqt_attn_tc_plan_destroy(p)
        print(json.dumps(dict(op="softmax_attention", tokens=n, heads=heads, dim_head=dh, ours_ms=ours, ours_tflops_algorithmic=fl / ours / 1e9,
                              ours_frac_burst=fl / ours / 1e9 / peaks["burst"], torch_ms=t, torch_tflops=fl / t["sdpa_bf16"] / 1e9,
                              speedup=t["sdpa_bf16"] / ours, max_rel_vs_sdpa_fp32=err, torch_timing=how)), flush=True)


def sweep_linear_attn():
    """LinearAttention core (imagen_pytorch3D.py:1001-1011): softmax over d of q (scaled), softmax over tokens of k, ctx = k^T v (64 x 64 per
    head), out = mish(q ctx).  4 N d^2 FLOP per head against 4 N d bf16 values of traffic: bandwidth / latency work, not tensor work."""
    heads, dh = 8, 64
    inner = heads * dh
    for n in (1728, 13824, 110592):
        qkvs = [torch.randn(n, 3 * inner, device=dev).bfloat16() for _ in range(3)]
        outs = [torch.empty(n, inner, dtype=torch.bfloat16, device=dev) for _ in range(3)]
        chunks = C.c_int(0)
        L.check(lib.diqt_linear_attention_chunks(n, C.byref(chunks)))
        stat = torch.empty(inner * 2, dtype=torch.float32, device=dev)
        part = torch.empty(chunks.value * heads * dh * dh, dtype=torch.float32, device=dev)
        nb = C.c_size_t(0)
        L.check(lib.diqt_linattn_tc_workspace_bytes(n, heads, C.byref(nb)))
        plans, keep = [], []
        for qkv, out in zip(qkvs, outs):
            ws = torch.empty(nb.value + 256, dtype=torch.uint8, device=dev)
            keep.append(ws)
            p, plan = qkv.data_ptr(), C.c_void_p(0)
            L.check(lib.diqt_linattn_tc_plan_create(p, p + inner * 2, p + 2 * inner * 2, 3 * inner, out.data_ptr(), inner, n, heads, dh ** -0.5, 1,
                                                    (ws.data_ptr() + 255) // 256 * 256, C.byref(plan)))
            plans.append(plan.value)

        def run(qkv, out):
            p = qkv.data_ptr()
            L.check(lib.diqt_linear_attention(p, p + inner * 2, p + 2 * inner * 2, 3 * inner, out.data_ptr(), inner, L.BF16, n, heads, dh, dh ** -0.5, 1,
                                              stat.data_ptr(), part.data_ptr(), L.current_stream()))

        reps = 20 if n < 50000 else 6
        ours, _ = timed([(lambda p=p: L.check(lib.diqt_linattn_tc_run(p, L.current_stream()))) for p in plans], reps)
        simt, _ = timed([(lambda q=q, o=o: run(q, o)) for q, o in zip(qkvs, outs)], reps)

        def ref(qkv):
            q, k, v = (t.view(n, heads, dh).permute(1, 0, 2) for t in qkv.chunk(3, dim=1))
            q = q.softmax(dim=-1) * dh ** -0.5
            k = k.softmax(dim=-2)
            ctx = torch.einsum("hnd,hne->hde", k, v)
            return F.mish(torch.einsum("hnd,hde->hne", q, ctx)).permute(1, 0, 2).reshape(n, inner)

        t, how = timed([(lambda q=q: ref(q)) for q in qkvs], reps)
        want = ref(qkvs[0].float())
        L.check(lib.diqt_linattn_tc_run(plans[0], L.current_stream())); torch.cuda.synchronize()
        err = ((outs[0].float() - want).abs().max() / want.abs().max()).item()
        for p in plans:
            lib.diqt_linattn_tc_plan_destroy(p)
        nbytes = 4.0 * n * inner * 2
        print(json.dumps(dict(op="linear_attention", tokens=n, heads=heads, dim_head=dh, ours_ms=ours, ours_gbs=nbytes / ours / 1e6,
                              ours_frac_hbm=nbytes / ours / 1e6 / peaks["hbm"], cuda_core_kernels_ms=simt, torch_ms={"bf16_eager": t}, speedup=t / ours,
                              max_rel_vs_torch_fp32=err, torch_timing=how,
                              note="tcgen05 kernels (csrc/linattn_tc.cu); algorithmic bytes = q, k, v read once + out written once; three rotating "
                                   "buffer sets")), flush=True)


if __name__ == "__main__":
    which = set(sys.argv[1:]) or {"conv", "norm", "attn", "linattn"}
    print(json.dumps(dict(info="cfg5 sweep", gpu=torch.cuda.get_device_name(0), torch=torch.__version__, cudnn=torch.backends.cudnn.version(),
                          peaks=peaks)), flush=True)
    if "conv" in which:
        sweep_conv()
    if "norm" in which:
        sweep_norm()
    if "attn" in which:
        sweep_attn()
    if "linattn" in which:
        sweep_linear_attn()
