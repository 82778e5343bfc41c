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


def infer_volume(sample_fn: Callable[[torch.Tensor], torch.Tensor], lowres_norm: torch.Tensor, *, patch: int, overlap: int,
                 raw_lowres: Optional[torch.Tensor] = None, batch_size: int = 1, fill_value: float = 0.0,
                 rank: int = 0, world: int = 1, group=None, batch_sample: bool = False) -> VolumeResult:
    """Denoise a whole (normalised) low-field volume patch by patch and stitch the result.

    sample_fn: (B, 1, P, P, P) low-field patches on the compute device -> (B, 1, P, P, P) denoised patches
               (e.g. `lambda lr: imagen.sample(batch_size=lr.shape[0], start_image_or_video=lr, start_at_unet_number=2)[0]`).
    lowres_norm: (X, Y, Z) normalised volume on the compute device.  raw_lowres: the un-normalised volume used by the
    5 % skip rule (defaults to `lowres_norm`).  fill_value: initial value of the output, `(0 - mean) / std` in the
    reference (test_all.py:211-212).
    With world > 1 every rank must call this with the same arguments; each denoises its block of the patch list and the
    patches are exchanged with ONE all_gather (torch.distributed; NCCL on GPUs, gloo in the CPU tests).
    """
    dev = lowres_norm.device
    raw = raw_lowres if raw_lowres is not None else lowres_norm
    grid = patch_grid(lowres_norm.shape, patch, overlap)
    kept = [o for o in grid if keep_patch(raw, o, patch)]
    start, stop, per = shard_range(len(kept), rank, world)
    mine = kept[start:stop]
    local = torch.zeros((per, patch, patch, patch), dtype=torch.float32, device=dev)
    for b0 in range(0, len(mine), batch_size):
        chunk = mine[b0:b0 + batch_size]
        lr = torch.stack([lowres_norm[i:i + patch, j:j + patch, k:k + patch] for (i, j, k) in chunk])[:, None].float().contiguous()
        out = sample_fn(lr)
        local[b0:b0 + len(chunk)] = out[:, 0].to(dev, torch.float32)
    if world > 1:
        import torch.distributed as dist
        gathered = torch.empty((world * per, patch, patch, patch), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, local, group=group)       # the only collective on the path
    else:
        gathered = local
    pred = torch.full(tuple(lowres_norm.shape), float(fill_value), dtype=torch.float32, device=dev)
    for n, origin in enumerate(kept):                                   # index order: later patches overwrite earlier ones
        r, slot = divmod(n, per)
        stitch_patch_(pred, gathered[r * per + slot], origin, patch, overlap, batch_sample)
    lo = lowres_norm.min()
    pred[lowres_norm == lo] = lo                                        # background mask (test_all.py:300)
    return VolumeResult(pred, len(kept), len(grid) - len(kept), per)
