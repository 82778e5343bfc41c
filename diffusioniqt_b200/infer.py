"""File-to-file inference: what /root/reference/test_all.py:182-315 does for one (low-field, high-field) pair of NIfTI volumes.

    load (nibabel get_fdata -> float32, :194-199) -> `cube` crop when the last side is not 256 (:86-95, 201-203) -> [0:256]^3 (:204-206)
    -> z-score with the dataset constants (:210-214) -> patches on the data.py:159-162 grid, 5 % skip rule -> `trainer.sample` per batch
    (:234) -> stitch with centre crops (:239-298) -> background mask (:300) -> NIfTI out with the high-field affine (:311-312)
    -> (MS-SSIM, PSNR) against the high-field volume (:316).

The sampler is the CUDA hot path (`Imagen.sample` / `ImagenTrainer.sample`); patches are cut and stitched on the device
(`volume.infer_volume`); with `world > 1` every rank denoises its block of the patch list (one all-gather).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import metrics as M
from . import volume as V
from .nifti import load_nifti, save_nifti


def cube(data):
    """test_all.py:86-95: the reference's fixed crop of the last two axes to 256 x 256."""
    if data.ndim > 3:
        return data[:, :, 2:258, 27:283]
    return data[2:258, 27:283]


@dataclass
class InferenceResult:
    prediction: torch.Tensor          # normalised, stitched volume (CPU)
    n_patches: int
    n_skipped: int
    ms_ssim: Optional[float] = None
    psnr: Optional[float] = None


def infer_nifti(sampler, configs, lr_path, out_path=None, hr_path=None, *, device="cuda", rank=0, world=1, group=None,
                evaluate_kernel_size: int = 11) -> InferenceResult:
    """sampler: an `ImagenTrainer` or `Imagen` of this package (anything with the reference's `.sample(...)` signature).
    configs: the reference's nested config dict (config/*.yaml): Data.mean / Data.std, Train.patch_size_sub, Train.batch_sample,
    Train.batch_sample_factor, Eval.overlap (the patch stride), Eval.batch_size."""
    mean, std = float(configs["Data"]["mean"]), float(configs["Data"]["std"])
    sub = int(configs["Train"]["patch_size_sub"])
    batch_sample = bool(configs["Train"].get("batch_sample", False))
    f = int(configs["Train"].get("batch_sample_factor", 3)) if batch_sample else 0
    patch = sub * f if batch_sample else sub                                   # data.py:147-150
    stride = int(configs["Eval"]["overlap"])
    batch_size = 1 if batch_sample else int(configs["Eval"].get("batch_size", 1))   # test_all.py:184-187

    raw, lr_affine, _ = load_nifti(lr_path)
    raw = raw.astype(np.float32)
    hr, affine = None, lr_affine
    if hr_path is not None:
        hr, affine, _ = load_nifti(hr_path)
        hr = hr.astype(np.float32)
    if raw.shape[-1] != 256 and raw.shape[-1] > 256:                           # :201-203 (only meaningful for the reference's scans)
        raw = cube(raw)
        hr = cube(hr) if hr is not None else None
    raw = raw[0:256, 0:256, 0:256]
    hr = hr[0:256, 0:256, 0:256] if hr is not None else None

    dev = torch.device(device)
    raw_t = torch.from_numpy(np.ascontiguousarray(raw)).to(dev)
    low = (raw_t - mean) / std

    def sample_fn(lr):
        return sampler.sample(batch_size=lr.shape[0], skip_steps=None, return_all_outputs=False, return_pil_images=False,
                              start_image_or_video=lr, start_at_unet_number=2)[0]

    res = V.infer_volume(sample_fn, low, patch=patch, overlap=stride, raw_lowres=raw_t, batch_size=batch_size, fill_value=(0.0 - mean) / std,
                         rank=rank, world=world, group=group, batch_sample=batch_sample, sub_f=f)
    pred = res.volume.cpu()
    out = InferenceResult(pred, res.n_patches, res.n_skipped)
    if out_path is not None and rank == 0:
        save_nifti(pred.numpy(), affine, out_path)                             # :311-312 (the normalised prediction, as in the reference)
    if hr is not None:
        hr_n = (torch.from_numpy(np.ascontiguousarray(hr)) - mean) / std       # :213
        out.ms_ssim, out.psnr = M.evaluate(hr_n, pred, kernel_size=evaluate_kernel_size)
    return out
