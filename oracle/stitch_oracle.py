"""CPU oracle for patch extraction / stitching  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A literal, loop-by-loop numpy restatement of the host-side code around the sampler in the reference's inference
script: patch grid and skip rule (/root/reference/data.py:159-162, 192-196), the stitch of denoised patches into the
volume (/root/reference/test_all.py:239-298, both the plain and the batch_sample branch) and the background mask (:300).
`test_all.py` is a script with hard-coded cluster paths and `data.py` needs nibabel, so neither can be imported: parity
for this part is UNPINNED by reference runs; the restatement follows the source line by line instead.
"""
from __future__ import annotations

import numpy as np


def patch_index_list(shape, patch, stride):
    idx = []
    for i in range(0, shape[0] - patch + 1, stride):            # data.py:159
        for j in range(0, shape[1] - patch + 1, stride):        # :160
            for k in range(0, shape[2] - patch + 1, stride):    # :161
                idx.append([i, j, k])
    return idx


def is_skipped(raw, idx, patch, ratio=0.05):
    blk = raw[idx[0]:idx[0] + patch, idx[1]:idx[1] + patch, idx[2]:idx[2] + patch]
    return (np.count_nonzero(blk) / float(patch * patch * patch)) < ratio          # data.py:192-196


def stitch(pred_ary, outputs, idxs, patch_size, overlap, batch_sample):
    """pred_ary: (X,Y,Z) array modified in place; outputs[n]: (P,P,P) denoised patch n; idxs[n]: its origin."""
    op = overlap // 2
    V = pred_ary.shape[-1]
    for out, idx in zip(outputs, idxs):
        if overlap < patch_size:
            if not batch_sample:
                # test_all.py:244-263 (per-axis face rules; see volume.crop_margins for the :243 caveat)
                ops = [op] * 6
                for a in range(3):
                    if idx[a] == 0:
                        ops[2 * a] = 0
                    if V - patch_size <= idx[a] + patch_size:
                        ops[2 * a + 1] = 0
            else:
                # test_all.py:270-293
                ops = [op] * 6
                for a in range(3):
                    if idx[a] == 0:
                        ops[2 * a] = 0
                    if (V == idx[a] + patch_size) or (V - patch_size <= idx[a]):
                        ops[2 * a + 1] = 0
            xs, xe, ys, ye, zs, ze = ops
            pred_ary[idx[0] + xs: idx[0] + patch_size - xe, idx[1] + ys: idx[1] + patch_size - ye, idx[2] + zs: idx[2] + patch_size - ze] = \
                out[xs:patch_size - xe, ys:patch_size - ye, zs:patch_size - ze]
        else:
            pred_ary[idx[0]: idx[0] + patch_size, idx[1]: idx[1] + patch_size, idx[2]: idx[2] + patch_size] = out   # :265, :298
    return pred_ary


def background_mask(pred_ary, lowres):
    min_val = lowres.min()
    pred_ary[np.where(lowres == min_val)] = min_val                                  # test_all.py:300
    return pred_ary
