"""Patch grid / sharding / stitching host logic, single process and world_size = 2 over gloo (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from diffusioniqt_b200 import volume as V
from oracle import stitch_oracle as so


def _volume(shape, seed=0, hole=True):
    rs = np.random.RandomState(seed)
    raw = rs.uniform(0, 900, shape).astype(np.float32)
    if hole:
        raw[: shape[0] // 3, : shape[1] // 2] = 0.0          # empty corner: exercises the 5 % skip rule and the background mask
    return raw


@pytest.mark.parametrize("shape,patch,stride", [((96, 96, 96), 32, 16), ((64, 64, 64), 32, 32), ((80, 96, 112), 32, 24)])
def test_grid_and_skip_rule_match_oracle(shape, patch, stride):
    raw = _volume(shape)
    grid = V.patch_grid(shape, patch, stride)
    assert [list(g) for g in grid] == so.patch_index_list(shape, patch, stride)
    for g in grid:
        assert V.keep_patch(torch.from_numpy(raw), g, patch) == (not so.is_skipped(raw, list(g), patch))


@pytest.mark.parametrize("batch_sample", [False, True])
@pytest.mark.parametrize("shape,patch,stride", [((96, 96, 96), 32, 16), ((64, 64, 64), 32, 32), ((128, 128, 128), 64, 32), ((96, 96, 96), 32, 48)])
def test_stitch_matches_oracle(shape, patch, stride, batch_sample):
    rs = np.random.RandomState(1)
    grid = V.patch_grid(shape, patch, stride)
    outs = [rs.standard_normal((patch,) * 3).astype(np.float32) for _ in grid]
    want = so.stitch(np.full(shape, -0.72, np.float32), outs, [list(g) for g in grid], patch, stride, batch_sample)
    got = torch.full(shape, -0.72)
    for o, g in zip(outs, grid):
        V.stitch_patch_(got, torch.from_numpy(o), g, patch, stride, batch_sample)
    assert np.array_equal(got.numpy(), want)


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 343, 344):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b, per = V.shard_range(n, r, world)
                assert b - a <= per
                seen += list(range(a, b))
            assert seen == list(range(n))


def _fake_sampler(lr):
    return lr * 0.5 + 1.0 + lr.flip(-1) * 0.25        # deterministic, patch-local


def _reference_volume(raw, mean, std, patch, stride):
    lr = (raw - mean) / std
    grid = so.patch_index_list(raw.shape, patch, stride)
    kept = [g for g in grid if not so.is_skipped(raw, g, patch)]
    outs = [_fake_sampler(torch.from_numpy(lr[g[0]:g[0] + patch, g[1]:g[1] + patch, g[2]:g[2] + patch].copy())[None, None])[0, 0].numpy() for g in kept]
    pred = np.full(raw.shape, (0 - mean) / std, np.float32)
    so.stitch(pred, outs, kept, patch, stride, False)
    return so.background_mask(pred, lr), len(kept)


def test_infer_volume_single_rank_matches_oracle():
    raw = _volume((96, 96, 96), seed=3)
    mean, std = 271.648, 377.117
    want, nkept = _reference_volume(raw, mean, std, 32, 16)
    res = V.infer_volume(_fake_sampler, torch.from_numpy((raw - mean) / std), patch=32, overlap=16, raw_lowres=torch.from_numpy(raw),
                         batch_size=5, fill_value=(0 - mean) / std)
    assert res.n_patches == nkept and res.n_skipped > 0
    assert np.allclose(res.volume.numpy(), want, atol=1e-6)


def _worker(rank, world, port, raw, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mean, std = 271.648, 377.117
    res = V.infer_volume(_fake_sampler, torch.from_numpy((raw - mean) / std), patch=32, overlap=16, raw_lowres=torch.from_numpy(raw),
                         batch_size=4, fill_value=(0 - mean) / std, rank=rank, world=world)
    ret[rank] = (res.volume.numpy(), res.n_patches, res.patches_per_rank)
    dist.barrier()
    dist.destroy_process_group()


def test_infer_volume_two_ranks_gloo():
    raw = _volume((96, 96, 96), seed=4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, raw, ret), nprocs=2, join=True)
    want, nkept = _reference_volume(raw, 271.648, 377.117, 32, 16)
    for r in (0, 1):
        vol, n, per = ret[r]
        assert n == nkept and per == (nkept + 1) // 2
        assert np.allclose(vol, want, atol=1e-6)


def test_sub_volume_retiling_matches_the_pinned_oracle():
    """volume.split_sub_volumes / merge_sub_volumes against the oracle helpers (which are pinned on utils_mine.py)."""
    from oracle import unet_oracle as uo
    x = torch.randn(2, 12, 12, 12)
    for f in (2, 3):
        sub = V.split_sub_volumes(x, f)
        for b in range(2):
            want = uo.split_sub_volumes(x[b][None, None], f)[:, 0]
            assert torch.equal(sub[b * f ** 3:(b + 1) * f ** 3], want)
        assert torch.equal(V.merge_sub_volumes(sub, f), x)


def test_infer_volume_batch_sample_layout_cpu():
    """sub_f = 3: every 12^3 patch travels as 27 sub-volumes of 4^3 (data.py:147-150, test_all.py:230-231, 267-268) and is stitched with
    the batch_sample face rule."""
    import numpy as np
    from oracle import stitch_oracle as so
    N, P, stride, f = 28, 12, 8, 3
    low = torch.randn(N, N, N)
    low[:5] = low.min()
    seen = []

    def fake(lr):
        assert lr.shape[1:] == (1, 4, 4, 4) and lr.shape[0] % 27 == 0
        seen.append(lr.shape[0])
        return lr * 2 + 1

    res = V.infer_volume(fake, low, patch=P, overlap=stride, raw_lowres=low - low.min(), fill_value=-3.0, batch_sample=True, sub_f=f)
    idxs = [i for i in so.patch_index_list(low.shape, P, stride) if not so.is_skipped((low - low.min()).numpy(), i, P)]
    outs = [(low[i:i + P, j:j + P, k:k + P] * 2 + 1).numpy() for i, j, k in idxs]
    want = so.background_mask(so.stitch(np.full((N, N, N), -3.0, np.float32), outs, idxs, P, stride, True), low.numpy())
    assert torch.equal(res.volume, torch.from_numpy(want)) and seen and all(s == 27 for s in seen)


class _FakeSampler:
    """Stands in for ImagenTrainer / Imagen on the CPU: same `.sample(...)` keywords as test_all.py:234, output = 2 * lr + 1."""

    def __init__(self):
        self.calls = []

    def sample(self, batch_size=1, skip_steps=None, return_all_outputs=False, return_pil_images=False, start_image_or_video=None,
               start_at_unet_number=1, **kw):
        assert start_at_unet_number == 2 and batch_size == start_image_or_video.shape[0]
        self.calls.append(tuple(start_image_or_video.shape))
        return start_image_or_video * 2 + 1, [], []


@pytest.mark.parametrize("batch_sample", [False, True])
def test_infer_nifti_file_to_file_on_cpu(tmp_path, batch_sample):
    """The host side of test_all.py:182-316 (NIfTI in, z-score, patch grid / skip rule, stitch, mask, NIfTI out, metrics) with a stand-in
    sampler, against the loop-by-loop oracle; `batch_sample` routes every 24^3 patch through 27 sub-volumes of 8^3."""
    import numpy as np
    from diffusioniqt_b200.infer import infer_nifti
    from diffusioniqt_b200.nifti import load_nifti, save_nifti
    from oracle import stitch_oracle as so
    N, sub, stride = 48, (8 if batch_sample else 16), 8      # 48: the smallest side with five MS-SSIM scales of a 3-wide window
    P = sub * 3 if batch_sample else sub
    mean, std = 271.648, 377.117
    cfg = {"Data": {"norm": "z-score", "mean": mean, "std": std},
           "Train": {"batch_sample": batch_sample, "batch_sample_factor": 3, "patch_size_sub": sub}, "Eval": {"overlap": stride, "batch_size": 5}}
    rs = np.random.RandomState(3)
    raw = np.abs(rs.standard_normal((N, N, N)).astype(np.float32)) * 300
    raw[:P + 2] = 0.0                                            # air: skipped patches
    affine = np.diag([2.0, 2.0, 2.0, 1.0])
    save_nifti(raw, affine, tmp_path / "lr.nii.gz")
    save_nifti(raw * 1.05 + 3, affine, tmp_path / "hr.nii")
    fake = _FakeSampler()
    res = infer_nifti(fake, cfg, tmp_path / "lr.nii.gz", tmp_path / "out.nii.gz", tmp_path / "hr.nii", device="cpu", evaluate_kernel_size=3)
    low = (raw - mean) / std
    idxs = [i for i in so.patch_index_list(raw.shape, P, stride) if not so.is_skipped(raw, i, P)]
    outs = [low[i:i + P, j:j + P, k:k + P] * 2 + 1 for i, j, k in idxs]
    want = so.background_mask(so.stitch(np.full(raw.shape, (0 - mean) / std, np.float32), outs, idxs, P, stride, batch_sample), low)
    assert res.n_patches == len(idxs) and res.n_skipped > 0
    assert np.allclose(res.prediction.numpy(), want, atol=1e-6)
    data, aff, _ = load_nifti(tmp_path / "out.nii.gz")
    assert np.array_equal(data.astype(np.float32), res.prediction.numpy()) and np.allclose(aff, affine)
    assert res.psnr is not None and res.ms_ssim is not None
    side = sub
    assert all(c[1:] == (1, side, side, side) for c in fake.calls)
    if batch_sample:
        assert all(c[0] == 27 for c in fake.calls)               # one 24^3 patch = 27 sub-volumes per call (test_all.py:184-185)
    else:
        assert max(c[0] for c in fake.calls) == 5                # Eval.batch_size patches per call


def test_infer_volume_edge_cases_cpu():
    """Empty and ragged inputs: every patch skipped (an all-air volume), a volume that holds exactly one patch, a ragged last batch."""
    P, stride = 8, 4
    calls = []

    def fake(lr):
        calls.append(lr.shape[0])
        return lr + 1

    air = torch.zeros(16, 16, 16)
    res = V.infer_volume(fake, (air - 1.0), patch=P, overlap=stride, raw_lowres=air, fill_value=-9.0)
    assert res.n_patches == 0 and res.n_skipped == 27 and not calls
    assert torch.equal(res.volume, torch.full((16, 16, 16), -1.0))          # background mask: every voxel equals the minimum
    one = torch.rand(P, P, P) + 0.5
    res = V.infer_volume(fake, one, patch=P, overlap=stride, fill_value=0.0)
    assert res.n_patches == 1 and calls == [1]
    want = one + 1
    want[one == one.min()] = one.min()
    assert torch.equal(res.volume, want)
    calls.clear()
    vol = torch.rand(16, 16, 16) + 0.5
    res = V.infer_volume(fake, vol, patch=P, overlap=stride, batch_size=4)
    assert res.n_patches == 27 and calls == [4] * 6 + [3]
    assert V.shard_range(0, 1, 4) == (0, 0, 0) and V.shard_range(5, 3, 4) == (5, 5, 2)
