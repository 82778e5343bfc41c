"""Per-kernel parity of the CUDA library against plain fp32 PyTorch on the CPU (the library the reference calls).

Tolerances (BASELINE.json north_star): 1e-5 relative in fp32 mode, 1e-2 relative in bf16 mode, where
"relative" = max |a-b| / max |b| for one kernel's output.
"""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_rel
from oracle.unet_oracle import pixel_shuffle3d, pixel_unshuffle3d

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 1e-2


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _conv_reference(x, w, b, mode):
    if mode == "k3":
        return F.conv3d(x, w, b, padding=1)
    if mode == "k1":
        return F.conv3d(x, w, b)
    if mode == "down":
        return F.conv3d(pixel_unshuffle3d(x), w, b)
    return pixel_shuffle3d(F.mish(F.conv3d(x, w, b)))


def _conv_weight(mode, c_in, c_out, seed):
    k = 3 if mode == "k3" else 1
    cin_w = c_in * 8 if mode == "down" else c_in
    fan = cin_w * k ** 3
    return _rand(c_out, cin_w, k, k, k, seed=seed, scale=fan ** -0.5), _rand(c_out, seed=seed + 1, scale=0.1)


CONV_SHAPES = [
    # mode, n, (d0,d1,d2), c_in, c_out
    ("k3", 1, (8, 8, 8), 16, 16),
    ("k3", 2, (4, 6, 10), 32, 48),
    ("k3", 1, (16, 16, 16), 64, 64),
    ("k3", 1, (8, 8, 8), 192, 128),
    ("k1", 2, (4, 4, 4), 128, 64),
    ("k1", 1, (8, 8, 8), 192, 128),
    ("down", 1, (8, 8, 8), 64, 128),
    ("down", 2, (4, 8, 16), 16, 32),
    ("up", 1, (4, 4, 4), 128, 512),
    ("up", 2, (2, 4, 8), 32, 128),
]


@pytest.mark.parametrize("mode,n,dims,c_in,c_out", CONV_SHAPES)
def test_conv_simt_fp32(mode, n, dims, c_in, c_out):
    from diffusioniqt_b200 import ops
    x = _rand(n, c_in, *dims, seed=1)
    w, b = _conv_weight(mode, c_in, c_out, 2)
    want = _conv_reference(x, w, b, mode)
    got = ops.conv3d(ops.to_channels_last(x.cuda()), w, b, mode=mode, impl="simt")
    assert max_rel(ops.from_channels_last(got).cpu(), want) < FP32_TOL


@pytest.mark.parametrize("mode,n,dims,c_in,c_out", CONV_SHAPES)
def test_conv_simt_bf16(mode, n, dims, c_in, c_out):
    from diffusioniqt_b200 import ops
    x = _rand(n, c_in, *dims, seed=3).bfloat16().float()
    w, b = _conv_weight(mode, c_in, c_out, 4)
    want = _conv_reference(x, w.bfloat16().float(), b, mode)
    got = ops.conv3d(ops.to_channels_last(x.cuda(), torch.bfloat16), w, b, mode=mode, impl="simt")
    assert max_rel(ops.from_channels_last(got).cpu(), want) < BF16_TOL


TC_SHAPES = [
    ("k3", 1, (16, 16, 16), 64, 64),
    ("k3", 1, (8, 8, 8), 128, 128),
    ("k3", 1, (8, 8, 8), 192, 128),
    ("k3", 2, (4, 4, 4), 128, 128),      # volume smaller than the 8x4x4 box: batch folded into the tile
    ("k3", 1, (12, 12, 12), 64, 64),     # edge tiles (12 is not a multiple of the box)
    ("k3", 1, (32, 32, 32), 64, 64),
    ("k3", 3, (2, 2, 2), 128, 128),
    ("k1", 1, (8, 8, 8), 192, 128),
    ("k1", 1, (16, 16, 16), 128, 64),
    ("k1", 2, (4, 4, 4), 128, 256),
    ("down", 1, (16, 16, 16), 64, 64),
    ("down", 1, (8, 8, 8), 64, 128),
    ("up", 1, (4, 4, 4), 256, 1024),
    ("up", 1, (8, 8, 8), 128, 512),
]


@pytest.mark.parametrize("mode,n,dims,c_in,c_out", TC_SHAPES)
def test_conv_tcgen05_bf16(mode, n, dims, c_in, c_out):
    from diffusioniqt_b200 import ops
    x = _rand(n, c_in, *dims, seed=5).bfloat16().float()
    w, b = _conv_weight(mode, c_in, c_out, 6)
    want = _conv_reference(x, w.bfloat16().float(), b, mode)
    got = ops.conv3d(ops.to_channels_last(x.cuda(), torch.bfloat16), w, b, mode=mode, impl="tc")
    assert max_rel(ops.from_channels_last(got).cpu(), want) < BF16_TOL


ZM_SHAPES = [
    # n, (d0, d1, d2), c_in, c_out: z-march kernel needs d1 % 16 == 0, d2 % 8 == 0, channels in multiples of 64
    (1, (16, 16, 16), 64, 64),
    (1, (2, 16, 8), 64, 64),
    (1, (5, 16, 8), 64, 64),        # odd depth: last z-segment shorter
    (2, (8, 32, 16), 64, 64),
    (3, (4, 16, 24), 64, 64),       # more items than one slot pair per column, batch > 1
    (1, (32, 32, 32), 64, 64),
    (1, (64, 64, 64), 64, 64),      # BASELINE config 2 shape
    (1, (16, 16, 16), 128, 128),    # two K chunks, two output-channel groups (level 2 of the driver U-Net)
    (1, (32, 32, 32), 192, 128),    # ups.0.1.block1: concat input, three K chunks
    (1, (16, 32, 16), 128, 64),     # ups.1.1.block1 at a test-sized volume
    (2, (3, 16, 8), 64, 192),       # odd item count per channel group: one slot of the last pair idles
    (1, (8, 16, 16), 256, 256),     # mid_block width (deep_feature)
    (5, (2, 16, 8), 64, 128),       # several volumes per CTA slot: per-(volume, group) statistics flushes
    (1, (8, 16, 16), 512, 512),     # BASELINE config 5's widest layer: eight K chunks, eight output-channel groups
    (1, (4, 16, 8), 320, 448),      # widths that are multiples of 64 but not powers of two
]


@pytest.mark.parametrize("n,dims,c_in,c_out", ZM_SHAPES)
def test_conv_zmarch_bf16(n, dims, c_in, c_out):
    from diffusioniqt_b200 import ops
    x = _rand(n, c_in, *dims, seed=21).bfloat16().float()
    w, b = _conv_weight("k3", c_in, c_out, 22)
    want = _conv_reference(x, w.bfloat16().float(), b, "k3")
    got, stats = ops.conv3d(ops.to_channels_last(x.cuda(), torch.bfloat16), w, b, mode="k3", impl="zm", with_stats=True)
    stored = ops.from_channels_last(got).cpu()
    assert max_rel(stored, want) < BF16_TOL
    # fused statistics describe the stored (bf16-rounded) output exactly up to fp32 summation order
    assert not torch.isnan(stats).any()
    s = stats.sum(dim=1).cpu()
    assert max_rel(s[..., 0], stored.sum(dim=(2, 3, 4))) < 2e-4
    assert max_rel(s[..., 1], (stored ** 2).sum(dim=(2, 3, 4))) < 2e-4


PAIR_SHAPES = [
    # n, dims, c_in, c_out: shapes the z-march kernel runs as CTA pairs (even number of 8-wide x tiles, >= 2 column pairs)
    (2, (8, 32, 16), 64, 64),
    (1, (32, 32, 32), 64, 64),
    (1, (32, 32, 32), 192, 128),
    (1, (5, 16, 48), 64, 128),      # three column pairs: the second slot of the last pair idles; odd depth
    (3, (4, 32, 16), 128, 64),      # several volumes, two K chunks
    (1, (64, 64, 64), 64, 64),      # BASELINE config 2 shape
]


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("n,dims,c_in,c_out", PAIR_SHAPES)
def test_conv_zmarch_cta_pair_equals_single_cta(n, dims, c_in, c_out, fused, monkeypatch):
    monkeypatch.setenv("DIQT_ZM_2CTA_MIN_PAIRS", "2")     # by default only volumes from 64^3 up run as CTA pairs (where it pays)
    _pair_vs_single(n, dims, c_in, c_out, fused)


def _pair_vs_single(n, dims, c_in, c_out, fused):
    """tcgen05.mma.cta_group::2 version of the z-march kernel (two CTAs share every weight stage, M = 256 per MMA) against the single-CTA
    version: same products, same accumulation order per output element -> the same bits, statistics included; and against fp32 PyTorch."""
    from diffusioniqt_b200 import ops
    x = (_rand(n, c_in, *dims, seed=51) * 1.2 + 0.1).bfloat16().float()
    w, b = _conv_weight("k3", c_in, c_out, 52)
    gamma, beta = _rand(c_in, seed=53) * 0.2 + 1.0, _rand(c_in, seed=54) * 0.1
    ss = _rand(n, 2 * c_in, seed=55) * 0.3
    gn = dict(groups=8, gamma=gamma, beta=beta, scale_shift=ss, nblk=16, affine=n > 2) if fused else None
    xg = ops.to_channels_last(x.cuda(), torch.bfloat16)
    one, st1 = ops.conv3d(xg, w, b, mode="k3", impl="zm", with_stats=True, gn=gn, pair=False)
    two, st2 = ops.conv3d(xg, w, b, mode="k3", impl="zm", with_stats=True, gn=gn, pair=True)
    assert torch.equal(one, two)
    # the statistics rows are laid out per CTA slot; their totals must agree exactly up to fp32 summation order
    assert max_rel(st2.sum(dim=1), st1.sum(dim=1)) < 1e-5
    y = x
    if fused:
        y = F.group_norm(x, 8, gamma, beta, 1e-5) * (ss[:, :c_in, None, None, None] + 1) + ss[:, c_in:, None, None, None]
        y = F.mish(y)
    want = F.conv3d(y, w.bfloat16().float(), b, padding=1)
    assert max_rel(ops.from_channels_last(two).cpu(), want) < BF16_TOL


GN_FUSED_SHAPES = [
    # n, dims, c_in, c_out, film
    (1, (16, 16, 16), 64, 64, True),
    (1, (5, 16, 8), 64, 64, False),      # one column, odd depth
    (2, (8, 32, 16), 64, 128, True),     # two volumes: per-volume statistics and FiLM rows, two output-channel groups
    (1, (32, 32, 32), 192, 128, True),   # ups.0.1.block1: concat input, GroupNorm groups of 24 channels straddling the 64-channel chunks
    (1, (16, 16, 16), 128, 128, False),
    (1, (8, 16, 16), 256, 64, True),     # widest fusable input
    (1, (64, 64, 64), 64, 64, True),     # BASELINE config 2 shape
]


@pytest.mark.parametrize("n,dims,c_in,c_out,film", GN_FUSED_SHAPES)
def test_conv_zmarch_fused_groupnorm_film_mish(n, dims, c_in, c_out, film):
    """Block.forward (GroupNorm -> FiLM -> Mish -> conv, imagen_pytorch3D.py:555-565) as ONE conv launch: the normalisation rides on the
    conv's load path.  Checked against fp32 PyTorch on the CPU, and bit for bit against the two-kernel path of this library (the
    transform produces the same bf16 operands the separate apply kernel would have stored)."""
    from diffusioniqt_b200 import ops
    x = (_rand(n, c_in, *dims, seed=31) * 1.7 + 0.3).bfloat16().float()
    w, b = _conv_weight("k3", c_in, c_out, 32)
    gamma, beta = _rand(c_in, seed=33) * 0.2 + 1.0, _rand(c_in, seed=34) * 0.1
    ss = _rand(n, 2 * c_in, seed=35) * 0.3 if film else None
    y = F.group_norm(x, 8, gamma, beta, 1e-5)
    if film:
        y = y * (ss[:, :c_in, None, None, None] + 1) + ss[:, c_in:, None, None, None]
    want = F.conv3d(F.mish(y), w.bfloat16().float(), b, padding=1)
    xg = ops.to_channels_last(x.cuda(), torch.bfloat16)
    got, stats = ops.conv3d(xg, w, b, mode="k3", impl="zm", with_stats=True, gn=dict(groups=8, gamma=gamma, beta=beta, scale_shift=ss, nblk=16))
    assert max_rel(ops.from_channels_last(got).cpu(), want) < BF16_TOL
    a = ops.group_norm_film_mish(xg, 8, gamma, beta, ss, nblk=16, grouped=True)
    two, stats2 = ops.conv3d(a, w, b, mode="k3", impl="zm", with_stats=True)
    assert torch.equal(got, two)
    assert torch.equal(stats, stats2)


@pytest.mark.parametrize("n,dims,c_in,c_out,film", [(4, (8, 16, 16), 64, 64, True), (3, (4, 16, 8), 320, 128, False), (5, (2, 16, 8), 128, 64, True)])
def test_conv_zmarch_fused_affine_mish_any_batch(n, dims, c_in, c_out, film):
    """Larger batches (BASELINE config 4) / wider layers: the statistics are finalised by diqt_gn_finalize and the conv applies the
    per-(volume, channel) affine + Mish on its load path; same bits as the separate apply kernel."""
    from diffusioniqt_b200 import ops
    x = (_rand(n, c_in, *dims, seed=41) * 1.3 - 0.2).bfloat16().float()
    w, b = _conv_weight("k3", c_in, c_out, 42)
    gamma, beta = _rand(c_in, seed=43) * 0.2 + 1.0, _rand(c_in, seed=44) * 0.1
    ss = _rand(n, 2 * c_in, seed=45) * 0.3 if film else None
    y = F.group_norm(x, 8, gamma, beta, 1e-5)
    if film:
        y = y * (ss[:, :c_in, None, None, None] + 1) + ss[:, c_in:, None, None, None]
    want = F.conv3d(F.mish(y), w.bfloat16().float(), b, padding=1)
    xg = ops.to_channels_last(x.cuda(), torch.bfloat16)
    got = ops.conv3d(xg, w, b, mode="k3", impl="zm", gn=dict(groups=8, gamma=gamma, beta=beta, scale_shift=ss, nblk=8, affine=True))
    assert max_rel(ops.from_channels_last(got).cpu(), want) < BF16_TOL
    a = ops.group_norm_film_mish(xg, 8, gamma, beta, ss, nblk=8, grouped=False)
    assert torch.equal(got, ops.conv3d(a, w, b, mode="k3", impl="zm"))


@pytest.mark.parametrize("mode,n,dims,c_in,c_out", [("k3", 1, (16, 16, 16), 64, 64), ("k3", 2, (8, 8, 8), 128, 128), ("down", 1, (16, 16, 16), 64, 128),
                                                     ("k1", 2, (8, 8, 8), 128, 256), ("k3", 1, (12, 12, 12), 64, 64)])
def test_conv_tcgen05_fused_stats(mode, n, dims, c_in, c_out):
    from diffusioniqt_b200 import ops
    x = _rand(n, c_in, *dims, seed=23).bfloat16().float()
    w, b = _conv_weight(mode, c_in, c_out, 24)
    got, stats = ops.conv3d(ops.to_channels_last(x.cuda(), torch.bfloat16), w, b, mode=mode, impl="tc", with_stats=True)
    assert stats is not None
    stored = ops.from_channels_last(got).cpu()
    s = stats.sum(dim=1).cpu()
    assert max_rel(s[..., 0], stored.sum(dim=(2, 3, 4))) < 2e-4
    assert max_rel(s[..., 1], (stored ** 2).sum(dim=(2, 3, 4))) < 2e-4


def test_conv_zmarch_equals_per_tap_kernel():
    """Same operands, same fp32 accumulation, different summation order only."""
    from diffusioniqt_b200 import ops
    x = ops.to_channels_last(_rand(1, 64, 16, 32, 16, seed=25).cuda(), torch.bfloat16)
    w, b = _conv_weight("k3", 64, 64, 26)
    a = ops.conv3d(x, w, b, mode="k3", impl="zm").float()
    c = ops.conv3d(x, w, b, mode="k3", impl="tc").float()
    assert max_rel(a, c) < 2 ** -7


def test_conv_tcgen05_matches_simt_closely():
    """Same bf16 operands, fp32 accumulation in both: the two kernel families agree to bf16 output rounding."""
    from diffusioniqt_b200 import ops
    x = ops.to_channels_last(_rand(1, 64, 16, 16, 16, seed=7).cuda(), torch.bfloat16)
    w, b = _conv_weight("k3", 64, 64, 8)
    a = ops.conv3d(x, w, b, mode="k3", impl="tc").float()
    c = ops.conv3d(x, w, b, mode="k3", impl="simt").float()
    assert max_rel(a, c) < 2 ** -7


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("n,dims,c,groups,film", [(1, (8, 8, 8), 64, 8, True), (2, (4, 6, 10), 96, 8, False), (3, (2, 2, 2), 128, 8, True),
                                                   (1, (16, 16, 16), 192, 8, True)])
@pytest.mark.parametrize("grouped", [False, True])
def test_group_norm_film_mish(dtype, tol, n, dims, c, groups, film, grouped):
    from diffusioniqt_b200 import ops
    x = (_rand(n, c, *dims, seed=9) * 1.7 + 0.4)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    gamma, beta = 1 + 0.2 * _rand(c, seed=10), 0.1 * _rand(c, seed=11)
    ss = 0.5 * _rand(n, 2 * c, seed=12) if film else None
    want = F.group_norm(x, groups, gamma, beta, eps=1e-5)
    if film:
        sc, sh = ss[:, :c, None, None, None], ss[:, c:, None, None, None]
        want = want * (sc + 1) + sh
    want = F.mish(want)
    got = ops.group_norm_film_mish(ops.to_channels_last(x.cuda(), dtype), groups, gamma, beta, ss.cuda() if film else None, nblk=5, grouped=grouped)
    assert max_rel(ops.from_channels_last(got).cpu(), want) < tol


@pytest.mark.parametrize("grouped", [False, True])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_se_scale_residual(dtype, tol, grouped):
    from diffusioniqt_b200 import ops
    n, c, dims = 2, 64, (6, 4, 8)
    h, r = _rand(n, c, *dims, seed=13), _rand(n, c, *dims, seed=14)
    if dtype == torch.bfloat16:
        h, r = h.bfloat16().float(), r.bfloat16().float()
    w1, w2 = _rand(c // 16, c, seed=15, scale=0.3), _rand(c, c // 16, seed=16, scale=0.8)
    y = torch.sigmoid(F.linear(torch.relu(F.linear(h.mean(dim=(2, 3, 4)), w1)), w2))
    want = h * y[:, :, None, None, None] + r
    out, gate, part = ops.se_scale_residual(ops.to_channels_last(h.cuda(), dtype), ops.to_channels_last(r.cuda(), dtype), w1, w2, nblk=37 if grouped else 7,
                                            grouped=grouped)
    if not grouped:
        assert max_rel(gate.cpu(), y) < 1e-5
    assert max_rel(ops.from_channels_last(out).cpu(), want) < tol
    # the fused statistics describe the stored output
    stored = ops.from_channels_last(out).cpu()
    s = part.sum(dim=1).cpu()
    assert max_rel(s[..., 0], stored.sum(dim=(2, 3, 4))) < 1e-4
    assert max_rel(s[..., 1], (stored ** 2).sum(dim=(2, 3, 4))) < 1e-4


@pytest.mark.parametrize("grouped", [False, True])
@pytest.mark.parametrize("n,c,dims,nblk", [(1, 64, (32, 32, 32), 37), (2, 128, (16, 16, 24), 8), (1, 256, (8, 16, 16), 4), (1, 64, (17, 16, 16), 5),
                                            (1, 64, (64, 64, 64), 148)])
def test_se_scale_residual_ring_kernel(n, c, dims, nblk, grouped, monkeypatch):
    """The bulk-copy ring version of the SE / residual pass (bf16, contiguous rows, >= 256 voxels per CTA) against fp32 PyTorch, and
    against the register-staged kernel: the same products and roundings -> the same stored bits; statistics equal up to summation order.
    (17, 16, 16) with 5 CTAs leaves ragged last tiles."""
    import importlib
    from diffusioniqt_b200 import ops
    h, r = _rand(n, c, *dims, seed=81).bfloat16().float(), _rand(n, c, *dims, seed=82).bfloat16().float()
    w1, w2 = _rand(c // 16, c, seed=83, scale=0.3), _rand(c, c // 16, seed=84, scale=0.8)
    y = torch.sigmoid(F.linear(torch.relu(F.linear(h.mean(dim=(2, 3, 4)), w1)), w2))
    want = h * y[:, :, None, None, None] + r
    hg, rg = ops.to_channels_last(h.cuda(), torch.bfloat16), ops.to_channels_last(r.cuda(), torch.bfloat16)
    out, gate, part = ops.se_scale_residual(hg, rg, w1, w2, nblk=nblk, grouped=grouped)
    stored = ops.from_channels_last(out).cpu()
    assert max_rel(stored, want) < BF16_TOL
    s = part.sum(dim=1).cpu()
    assert max_rel(s[..., 0], stored.sum(dim=(2, 3, 4))) < 1e-4
    assert max_rel(s[..., 1], (stored ** 2).sum(dim=(2, 3, 4))) < 1e-4


def test_grouped_statistics_sum_to_the_partial_rows():
    """Group rows are sums of consecutive partial rows (fixed order); several group sizes incl. a ragged last group; tickets reset."""
    from diffusioniqt_b200 import ops
    x = ops.to_channels_last(_rand(2, 64, 8, 8, 8, seed=18).cuda(), torch.bfloat16)
    for nblk in (1, 7, 16, 37, 148):
        part, grp = ops.channel_stats_grouped(x, nblk)
        assert not torch.isnan(grp).any() and grp.shape[1] <= 16
        assert max_rel(grp.sum(dim=1).cpu(), part.sum(dim=1).cpu()) < 1e-6
        part2, grp2 = ops.channel_stats_grouped(x, nblk)
        assert torch.equal(grp, grp2)


def test_conv_grouped_statistics():
    """conv epilogue statistics through the grouped sink (z-march: two rows per CTA; per-tap kernel: one)."""
    import ctypes as C
    from diffusioniqt_b200 import lib as L, ops
    for impl, dims, cin, cout in (("zm", (16, 16, 16), 64, 128), ("zm", (64, 64, 64), 64, 64), ("tc", (8, 8, 8), 128, 128)):
        x = ops.to_channels_last(_rand(1, cin, *dims, seed=27).cuda(), torch.bfloat16)
        w, b = _conv_weight("k3", cin, cout, 28)
        got, stats = ops.conv3d(x, w, b, mode="k3", impl=impl, with_stats=True, grouped=True)
        part, grp = stats
        assert grp.shape[1] <= 16 and not torch.isnan(grp).any()
        assert max_rel(grp.sum(dim=1).cpu(), part.sum(dim=1).cpu()) < 1e-6
        stored = ops.from_channels_last(got).cpu()
        assert max_rel(grp.sum(dim=1)[..., 0].cpu(), stored.sum(dim=(2, 3, 4))) < 2e-4


def test_channel_stats_is_deterministic():
    from diffusioniqt_b200 import ops
    x = ops.to_channels_last(_rand(2, 64, 8, 8, 8, seed=17).cuda(), torch.bfloat16)
    a, b = ops.channel_stats(x, 13), ops.channel_stats(x, 13)
    torch.cuda.synchronize()
    assert torch.equal(a, b)


@pytest.mark.parametrize("n,dims,c_in,c_out", [(1, (8, 8, 8), 2, 64), (2, (4, 6, 10), 1, 64), (1, (16, 16, 16), 2, 128)])
def test_init_conv_im2col_tensor_core_path(n, dims, c_in, c_out):
    """init_conv (:1291) as im2col (K = 27 * c_in -> 64 bf16 columns) + 1x1x1 tcgen05 conv against F.conv3d on the bf16-rounded input."""
    import ctypes as C
    from diffusioniqt_b200 import lib as L, ops
    lib = L.load()
    x = _rand(n, c_in, *dims, seed=11)
    w, b = _rand(c_out, c_in, 3, 3, 3, seed=12, scale=(27 * c_in) ** -0.5), _rand(c_out, seed=13, scale=0.1)
    want = F.conv3d(x.bfloat16().float(), w.bfloat16().float(), b, padding=1)
    xc = x.cuda().contiguous()
    vox = dims[0] * dims[1] * dims[2]
    planes = (C.c_void_p * c_in)(*[xc.data_ptr() + ci * vox * 4 for ci in range(c_in)])
    strides = (C.c_int64 * c_in)(*[c_in * vox] * c_in)
    col = torch.empty(n, *dims, 64, dtype=torch.bfloat16, device="cuda")
    L.check(lib.diqt_init_im2col(planes, strides, c_in, col.data_ptr(), n, *dims, L.current_stream()), "init_im2col")
    # the im2col tensor itself: column k = tap * c_in + ci holds the tap-shifted, zero-padded plane
    xp = F.pad(x.bfloat16().float(), (1, 1, 1, 1, 1, 1))
    ref = torch.zeros(n, *dims, 64)
    for tap in range(27):
        kz, ky, kx = tap // 9, (tap // 3) % 3, tap % 3
        for ci in range(c_in):
            ref[..., tap * c_in + ci] = xp[:, ci, kz:kz + dims[0], ky:ky + dims[1], kx:kx + dims[2]]
    assert torch.equal(col.float().cpu(), ref)
    w2 = torch.zeros(c_out, 64)
    w2[:, :27 * c_in] = w.permute(0, 2, 3, 4, 1).reshape(c_out, 27 * c_in)
    got = ops.conv3d(col, w2.reshape(c_out, 64, 1, 1, 1), b, mode="k1", impl="tc")
    assert max_rel(ops.from_channels_last(got).cpu(), want) < BF16_TOL


@pytest.mark.parametrize("grouped", [False, True])
@pytest.mark.parametrize("n,dims,c_in,c_out", [(1, (8, 8, 16), 2, 64), (2, (3, 12, 32), 1, 64), (1, (16, 16, 16), 2, 128), (1, (5, 20, 64), 2, 64),
                                               (1, (64, 64, 64), 2, 64),
                                               # work items of 4 y rows (d2 = 128; c_out = 128 at d2 = 64: two accumulator sets of 256 columns),
                                               # three volumes through one CTA's statistics, more items than CTAs with a ragged last y tile
                                               (1, (4, 8, 128), 2, 64), (1, (6, 12, 64), 2, 128), (3, (2, 8, 32), 2, 64), (1, (40, 36, 64), 1, 64)])
def test_init_conv_fused_tensor_core_kernel(n, dims, c_in, c_out, grouped):
    """init_conv (:1291) as ONE kernel: im2col rows built in shared memory, tcgen05 GEMM, bias, bf16 store and the channel statistics the first
    GroupNorm needs -- against F.conv3d on the bf16-rounded input, bit for bit against the round-1 two-kernel path, statistics against
    the stored output.  d1 = 12 and 20 are not multiples of the 8-row work item: the last item of a plane is partly outside the volume."""
    import ctypes as C
    from diffusioniqt_b200 import lib as L, ops
    lib = L.load()
    assert lib.diqt_init_conv_tc_supported(c_in, c_out, dims[1], dims[2])
    x = _rand(n, c_in, *dims, seed=61)
    w, b = _rand(c_out, c_in, 3, 3, 3, seed=62, scale=(27 * c_in) ** -0.5), _rand(c_out, seed=63, scale=0.1)
    want = F.conv3d(x.bfloat16().float(), w.bfloat16().float(), b, padding=1)
    xc = x.cuda().contiguous()
    vox = dims[0] * dims[1] * dims[2]
    planes = (C.c_void_p * c_in)(*[xc.data_ptr() + ci * vox * 4 for ci in range(c_in)])
    strides = (C.c_int64 * c_in)(*[c_in * vox] * c_in)
    w2 = torch.zeros(c_out, 64)
    w2[:, :27 * c_in] = w.permute(0, 2, 3, 4, 1).reshape(c_out, 27 * c_in)
    rows, chunks = torch.arange(c_out)[:, None], torch.arange(8)[None, :]
    wsw = torch.gather(w2.bfloat16().view(c_out, 8, 8), 1, (chunks ^ (rows & 7))[:, :, None].expand(c_out, 8, 8)).contiguous().cuda()
    bias = b.cuda()
    nb = C.c_int(0)
    L.check(lib.diqt_init_conv_tc_blocks(n, dims[0], dims[1], C.byref(nb)))
    out = torch.full((n, *dims, c_out), float("nan"), dtype=torch.bfloat16, device="cuda")
    part = torch.full((n, nb.value, c_out, 2), float("nan"), device="cuda")
    ng = C.c_int(0)
    L.check(lib.diqt_stats_groups(nb.value, 1, C.byref(ng)))
    grp = torch.full((n, ng.value, c_out, 2), float("nan"), device="cuda")
    tick = torch.zeros(16 * n, dtype=torch.int32, device="cuda")
    L.check(lib.diqt_init_conv_tc(planes, strides, c_in, wsw.data_ptr(), bias.data_ptr(), out.data_ptr(), c_out, n, *dims, c_out, part.data_ptr(),
                                  grp.data_ptr() if grouped else 0, tick.data_ptr() if grouped else 0, L.current_stream()), "init_conv_tc")
    torch.cuda.synchronize()
    stored = ops.from_channels_last(out).cpu()
    assert max_rel(stored, want) < BF16_TOL
    assert not torch.isnan(part).any()
    s = part.sum(dim=1).cpu()
    assert max_rel(s[..., 0], stored.sum(dim=(2, 3, 4))) < 2e-4
    assert max_rel(s[..., 1], (stored ** 2).sum(dim=(2, 3, 4))) < 2e-4
    if grouped:
        assert int(tick.abs().sum()) == 0
        assert max_rel(grp.sum(dim=1).cpu(), s) < 1e-5
    # the round-1 path: im2col rows in global memory + the 1x1x1 tcgen05 conv (same products, same accumulation order)
    col = torch.empty(n, *dims, 64, dtype=torch.bfloat16, device="cuda")
    L.check(lib.diqt_init_im2col(planes, strides, c_in, col.data_ptr(), n, *dims, L.current_stream()), "init_im2col")
    two = ops.conv3d(col, w2.reshape(c_out, 64, 1, 1, 1), b, mode="k1", impl="tc")
    assert torch.equal(two, out)


@pytest.mark.parametrize("n,dims,c_in,c_out", [(1, (16, 16, 16), 128, 128), (1, (16, 16, 16), 64, 64), (2, (8, 8, 8), 128, 128), (1, (8, 8, 8), 256, 256),
                                               (1, (16, 16, 16), 512, 512), (1, (4, 6, 10), 64, 128), (1, (16, 16, 16), 192, 128)])
def test_conv_tcgen05_split_k_small_volumes(n, dims, c_in, c_out):
    """Small volumes (fewer 128-voxel tiles than half the SMs) run the per-tap kernel with the K-blocks of a tile dealt to several CTAs; the last
    CTA to finish sums the fp32 partials in split order.  Against fp32 PyTorch, against the unsplit kernel (same products, another
    summation tree: equal up to bf16 rounding of a few elements), repeatable bit for bit, statistics consistent with the stored output."""
    from diffusioniqt_b200 import ops
    x = _rand(n, c_in, *dims, seed=71).bfloat16().float()
    w, b = _conv_weight("k3", c_in, c_out, 72)
    want = _conv_reference(x, w.bfloat16().float(), b, "k3")
    xg = ops.to_channels_last(x.cuda(), torch.bfloat16)
    got, stats = ops.conv3d(xg, w, b, mode="k3", impl="tc", with_stats=True)
    again, _ = ops.conv3d(xg, w, b, mode="k3", impl="tc", with_stats=True)
    plain = ops.conv3d(xg, w, b, mode="k3", impl="tc", split_k=False)
    stored = ops.from_channels_last(got).cpu()
    assert max_rel(stored, want) < BF16_TOL
    assert torch.equal(got, again)
    assert max_rel(got.float().cpu(), plain.float().cpu()) < 8e-3
    if stats is not None:
        assert not torch.isnan(stats).any()
        s = stats.sum(dim=1).cpu()
        assert max_rel(s[..., 0], stored.sum(dim=(2, 3, 4))) < 2e-4
        assert max_rel(s[..., 1], (stored ** 2).sum(dim=(2, 3, 4))) < 2e-4
    # auto dispatch takes this path for such shapes
    auto = ops.conv3d(xg, w, b, mode="k3", impl="auto")
    assert torch.equal(auto, got)
