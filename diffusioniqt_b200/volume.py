"""Whole-volume inference around the sampler: patch grid, sharding over ranks, all-gather, stitching.

Mirrors what the reference's inference script does on the host around `trainer.sample`
(/root/reference/data.py:139-202 `supervisedIQT_INF`, /root/reference/test_all.py:182-300):
  * patch origins on a regular grid `range(0, N - P + 1, stride)` per axis, k fastest (data.py:159-162; the config
    key called `overlap` is the stride);
  * a patch is skipped when fewer than 5 % of its raw voxels are non-zero (data.py:153, 192-196);
  * every denoised patch writes its centre crop (margin `overlap // 2`, no margin on volume faces) into the output
    volume, patches applied in index order so later ones overwrite earlier ones (test_all.py:239-298);
  * voxels where the low-field input equals its minimum get that minimum (background mask, test_all.py:300).

Patches are independent, so ranks take contiguous blocks of the patch list and the only collective is one
all-gather of the denoised patches (SURVEY.md section 8 e).  Device memory movement (slicing, gather) uses torch;
the denoising itself is `Imagen.sample` (CUDA kernels).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def patch_grid(shape: Sequence[int], patch: int, stride: int) -> List[Tuple[int, int, int]]:
    """Patch origins in the reference's nested-loop order (data.py:159-162)."""
    return [(i, j, k)
            for i in range(0, shape[0] - patch + 1, stride)
            for j in range(0, shape[1] - patch + 1, stride)
            for k in range(0, shape[2] - patch + 1, stride)]


def keep_patch(raw_lowres: torch.Tensor, origin, patch: int, ratio: float = 0.05) -> bool:
    """data.py:192-196: skip when the non-zero proportion of the RAW (un-normalised) patch is below `ratio`."""
    i, j, k = origin
    blk = raw_lowres[i:i + patch, j:j + patch, k:k + patch]
    return (torch.count_nonzero(blk).item() / float(patch ** 3)) >= ratio


def crop_margins(origin, patch: int, vol: int, overlap: int, batch_sample: bool = False):
    """Per-axis (start, end) margins of the crop a patch writes (test_all.py:244-263 and :270-293).

    `vol` is the reference's `pred_ary.shape[-1]` (it uses the last dimension for every axis).  The margin is
    `overlap // 2` except: start = 0 for a patch that begins at the volume face; end = 0 for a patch near the far face -
    `vol - patch <= origin + patch` in the plain branch (:249, :255, :261), `vol == origin + patch or vol - patch <= origin`
    in the batch_sample branch (:275, :281, :287).  (The plain branch's outer test at :243 compares whole index vectors and
    raises for interior patches in the reference; the per-axis rules below are its evident intent.)"""
    op = overlap // 2
    if overlap >= patch:                       # :264-265, :297-298: no cropping at all
        return [(0, 0)] * 3
    out = []
    for a in range(3):
        o = origin[a]
        start = 0 if o == 0 else op
        if batch_sample:
            end = 0 if (vol == o + patch or vol - patch <= o) else op
        else:
            end = 0 if (vol - patch <= o + patch) else op
        out.append((start, end))
    return out


def stitch_patch_(pred: torch.Tensor, out_patch: torch.Tensor, origin, patch: int, overlap: int, batch_sample: bool = False) -> None:
    """Write one denoised patch (patch^3) into `pred` in place, exactly like one iteration of test_all.py:239-298."""
    vol = pred.shape[-1]
    (xs, xe), (ys, ye), (zs, ze) = crop_margins(origin, patch, vol, overlap, batch_sample)
    i, j, k = origin
    pred[i + xs:i + patch - xe, j + ys:j + patch - ye, k + zs:k + patch - ze] = out_patch[xs:patch - xe, ys:patch - ye, zs:patch - ze]


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Contiguous block of ceil(n/world) items per rank: (start, stop, per_rank). Tail ranks may get fewer (or none)."""
    per = (n_items + world - 1) // world if n_items else 0
    start = min(n_items, rank * per)
    return start, min(n_items, start + per), per


@dataclass
class VolumeResult:
    volume: torch.Tensor            # stitched prediction, same shape as the input volume
    n_patches: int                  # patches denoised (after the skip rule)
    n_skipped: int
    patches_per_rank: int


def split_sub_volumes(patches: torch.Tensor, f: int) -> torch.Tensor:
    """(B, P, P, P) -> (B * f^3, P/f, P/f, P/f) in the order of convertVolume2subVolume (utils_mine.py:25-42): sub-volume
    b = b0 + f*b1 + f*f*b2 is block (b0, b1, b2) along the three spatial dims."""
    B, P = patches.shape[0], patches.shape[-1]
    h = P // f
    v = patches.reshape(B, f, h, f, h, f, h)              # B, b0, x, b1, y, b2, z
    return v.permute(0, 5, 3, 1, 2, 4, 6).reshape(B * f ** 3, h, h, h)


def merge_sub_volumes(sub: torch.Tensor, f: int) -> torch.Tensor:
    """Inverse of `split_sub_volumes` (merge_sub_volumes, utils_mine.py:44-67): (B * f^3, h, h, h) -> (B, f*h, f*h, f*h)."""
    h = sub.shape[-1]
    B = sub.shape[0] // f ** 3
    v = sub.reshape(B, f, f, f, h, h, h)                  # B, b2, b1, b0, x, y, z
    return v.permute(0, 3, 4, 2, 5, 1, 6).reshape(B, f * h, f * h, f * h)


def infer_volume(sample_fn: Callable[[torch.Tensor], torch.Tensor], lowres_norm: torch.Tensor, *, patch: int, overlap: int,
                 raw_lowres: Optional[torch.Tensor] = None, batch_size: int = 1, fill_value: float = 0.0,
                 rank: int = 0, world: int = 1, group=None, batch_sample: bool = False, sub_f: int = 0,
                 gather_fn: Optional[Callable[[torch.Tensor], Optional[torch.Tensor]]] = None) -> Optional[VolumeResult]:
    """Denoise a whole (normalised) low-field volume patch by patch and stitch the result.

    sample_fn: (B, 1, P, P, P) low-field patches on the compute device -> (B, 1, P, P, P) denoised patches
               (e.g. `lambda lr: imagen.sample(batch_size=lr.shape[0], start_image_or_video=lr, start_at_unet_number=2)[0]`).
               With sub_f = f > 1 (`Train.batch_sample`, data.py:147-150, test_all.py:230-231, 267-268) every P^3 patch is handed over as
               its f^3 sub-volumes, (B * f^3, 1, P/f, P/f, P/f), and comes back the same way.
    lowres_norm: (X, Y, Z) normalised volume on the compute device.  raw_lowres: the un-normalised volume used by the
    5 % skip rule (defaults to `lowres_norm`).  fill_value: initial value of the output, `(0 - mean) / std` in the
    reference (test_all.py:211-212).
    With world > 1 every rank must call this with the same arguments; each denoises its block of the patch list and the
    patches are exchanged with ONE all_gather (torch.distributed; NCCL on GPUs, gloo in the CPU tests).
    gather_fn replaces the collective: it receives this rank's padded shard (per, P^3) and returns all shards concatenated in rank order
    (world * per, P^3), or None to stop before the stitch (callers that exchange the shards themselves; the single-process tests of the
    multi-rank path).
    On a CUDA device the patches are cut and stitched by two kernels of libdiqt_b200 (`diqt_gather_patches`, `diqt_stitch_patches`);
    CPU tensors (the multi-process host-logic tests, with a stand-in `sample_fn`) take the equivalent torch slicing below.
    """
    dev = lowres_norm.device
    raw = raw_lowres if raw_lowres is not None else lowres_norm
    f = int(sub_f) if sub_f and sub_f > 1 else 1
    if patch % f:
        raise ValueError(f"patch {patch} is not a multiple of the sub-volume factor {f}")
    grid = patch_grid(lowres_norm.shape, patch, overlap)
    kept = [o for o in grid if keep_patch(raw, o, patch)]
    start, stop, per = shard_range(len(kept), rank, world)
    mine = kept[start:stop]
    on_gpu = dev.type == "cuda"
    vol32 = lowres_norm.float().contiguous()
    local = torch.zeros((per, patch ** 3), dtype=torch.float32, device=dev)
    if on_gpu:
        from . import lib as L
        lib = L.load()
        d0, d1, d2 = (int(v) for v in vol32.shape)
        origins = torch.tensor(mine if mine else [[0, 0, 0]], dtype=torch.int32, device=dev)
    for b0 in range(0, len(mine), batch_size):
        chunk = mine[b0:b0 + batch_size]
        nb = len(chunk)
        if on_gpu:
            lr = torch.empty((nb * f ** 3, 1) + (patch // f,) * 3, dtype=torch.float32, device=dev)
            L.check(lib.diqt_gather_patches(vol32.data_ptr(), d0, d1, d2, origins[b0:b0 + nb].data_ptr(), nb, patch, f, lr.data_ptr(),
                                            L.current_stream()), "gather_patches")
        else:
            lr = torch.stack([vol32[i:i + patch, j:j + patch, k:k + patch] for (i, j, k) in chunk])
            lr = (split_sub_volumes(lr, f) if f > 1 else lr)[:, None].contiguous()
        out = sample_fn(lr)
        local[b0:b0 + nb] = out.to(dev, torch.float32).reshape(nb, patch ** 3)     # same (sub-volume) layout as `lr`
    if gather_fn is not None:
        gathered = gather_fn(local)
        if gathered is None:
            return None
        assert tuple(gathered.shape) == (world * per, patch ** 3), (tuple(gathered.shape), world, per)
    elif world > 1:
        import torch.distributed as dist
        gathered = torch.empty((world * per, patch ** 3), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, local, group=group)       # the only collective on the path
    else:
        gathered = local
    pred = torch.full(tuple(lowres_norm.shape), float(fill_value), dtype=torch.float32, device=dev)
    lo = vol32.min()
    if on_gpu and kept:
        g = [len(range(0, int(s) - patch + 1, overlap)) for s in vol32.shape]
        slot = torch.full((g[0] * g[1] * g[2],), -1, dtype=torch.int32)
        index = {o: n for n, o in enumerate(grid)}
        for n, o in enumerate(kept):                                    # rank r holds kept[r*per : (r+1)*per] -> row n of `gathered`
            slot[index[o]] = n
        slot = slot.to(dev)
        L.check(lib.diqt_stitch_patches(gathered.data_ptr(), slot.data_ptr(), g[0], g[1], g[2], overlap, patch, overlap, 1 if batch_sample else 0, f,
                                        pred.data_ptr(), d0, d1, d2, vol32.data_ptr(), float(lo), L.current_stream()), "stitch_patches")
        return VolumeResult(pred, len(kept), len(grid) - len(kept), per)
    for n, origin in enumerate(kept):                                   # index order: later patches overwrite earlier ones
        p3 = gathered[n].reshape((f ** 3,) + (patch // f,) * 3) if f > 1 else gathered[n].reshape(patch, patch, patch)
        if f > 1:
            p3 = merge_sub_volumes(p3, f)[0]
        stitch_patch_(pred, p3, origin, patch, overlap, batch_sample)
    pred[vol32 == lo] = lo                                              # background mask (test_all.py:300)
    return VolumeResult(pred, len(kept), len(grid) - len(kept), per)
