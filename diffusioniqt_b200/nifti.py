"""Minimal NIfTI-1 reader / writer: the on-disk format either side of the sampling path.

The reference reads low-field / high-field volumes with `nibabel.load(f).get_fdata()` (data.py:158, test_all.py:194-199) and writes the
stitched prediction with `nib.save(nib.Nifti1Image(pred, affine), '...nii.gz')` (test_all.py:311-312).  nibabel is a third-party
dependency that is not vendored in the reference tree (requirements.txt pins nibabel==4.0.2) and is absent in this image, so this module
restates the published NIfTI-1 single-file layout (348-byte header, `n+1` magic, voxel data at `vox_offset`, Fortran order) for exactly
the calls the reference makes.  Parity with nibabel is therefore unpinned; the tests check the header layout against the NIfTI-1
standard's field offsets, round trips, scaling, endianness and gzip handling.

    data, affine, header = load_nifti(path)          # data: float64 array like get_fdata(), affine: (4, 4) float64
    save_nifti(array, affine, path)                  # like nib.save(nib.Nifti1Image(array, affine), path)
"""
from __future__ import annotations

import gzip
import struct
from typing import Dict, Tuple

import numpy as np

# NIfTI-1 datatype codes -> numpy dtypes (nifti1.h DT_*)
_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16, 768: np.uint32,
           1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def _open(path, mode):
    path = str(path)
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


def _quaternion_affine(b, c, d, qfac, pixdim, qoffset):
    """qform -> affine (nifti1.h "METHOD 2")."""
    a2 = 1.0 - (b * b + c * c + d * d)
    a = np.sqrt(a2) if a2 > 0 else 0.0
    if a2 <= 0:                                   # renormalise (b, c, d) like the reference C library
        n = np.sqrt(b * b + c * c + d * d)
        b, c, d = b / n, c / n, d / n
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    zooms = np.array([pixdim[1], pixdim[2], pixdim[3] * (qfac if qfac in (-1.0, 1.0) else 1.0)])
    aff = np.eye(4)
    aff[:3, :3] = R * zooms[None, :]
    aff[:3, 3] = qoffset
    return aff


def load_nifti(path) -> Tuple[np.ndarray, np.ndarray, Dict]:
    """Read a single-file NIfTI-1 image (.nii / .nii.gz).  Returns (data, affine, header): `data` is float64 with the stored
    scl_slope / scl_inter applied, shaped dim[1..ndim] - what `nibabel.load(path).get_fdata()` returns."""
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError(f"{path}: not a NIfTI-1 file (only {len(raw)} bytes)")
    endian = "<"
    if struct.unpack("<i", raw[:4])[0] != 348:
        if struct.unpack(">i", raw[:4])[0] != 348:
            raise ValueError(f"{path}: sizeof_hdr is not 348 in either byte order")
        endian = ">"
    magic = raw[344:348]
    if magic[:3] not in (b"n+1", b"ni1"):
        raise ValueError(f"{path}: bad NIfTI-1 magic {magic!r}")
    if magic[:3] == b"ni1":
        raise NotImplementedError(f"{path}: header / image pairs (.hdr + .img) are not read; the reference uses single .nii(.gz) files")
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(endian + "2h", raw[70:74])
    pixdim = struct.unpack(endian + "8f", raw[76:108])
    vox_offset, scl_slope, scl_inter = struct.unpack(endian + "3f", raw[108:120])
    qform_code, sform_code = struct.unpack(endian + "2h", raw[252:256])
    quatern = struct.unpack(endian + "3f", raw[256:268])
    qoffset = struct.unpack(endian + "3f", raw[268:280])
    srow = np.array(struct.unpack(endian + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    ndim = dim[0]
    if not 1 <= ndim <= 7:
        raise ValueError(f"{path}: dim[0] = {ndim}")
    if datatype not in _DTYPES:
        raise NotImplementedError(f"{path}: NIfTI datatype code {datatype} is not supported")
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    count = int(np.prod(shape))
    off = int(vox_offset) if vox_offset >= 352 else 352
    if len(raw) < off + count * dt.itemsize:
        raise ValueError(f"{path}: file holds {len(raw) - off} data bytes, header promises {count * dt.itemsize}")
    data = np.frombuffer(raw, dtype=dt, count=count, offset=off).reshape(shape, order="F").astype(np.float64)
    if np.isfinite(scl_slope) and scl_slope != 0.0 and not (scl_slope == 1.0 and scl_inter == 0.0):
        data = data * float(scl_slope) + (float(scl_inter) if np.isfinite(scl_inter) else 0.0)
    if sform_code > 0:                              # nibabel's get_best_affine order: sform, then qform, then pixdim
        affine = np.vstack([srow, [0.0, 0.0, 0.0, 1.0]])
    elif qform_code > 0:
        affine = _quaternion_affine(*quatern, pixdim[0], pixdim, qoffset)
    else:
        affine = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0]).astype(np.float64)
        affine[:3, 3] = -0.5 * (np.array(shape[:3] + (1,) * (3 - min(3, len(shape))), dtype=np.float64)[:3] - 1) * np.diag(affine)[:3]
    header = dict(dim=dim, datatype=datatype, bitpix=bitpix, pixdim=pixdim, vox_offset=vox_offset, scl_slope=scl_slope, scl_inter=scl_inter,
                  qform_code=qform_code, sform_code=sform_code, endian=endian, shape=shape)
    return data, affine, header


def save_nifti(array, affine, path) -> None:
    """Write `array` (up to 7-D; numpy or torch CPU tensor) as a single-file NIfTI-1 image with `affine` as the sform (code 2,
    'aligned'), the way `nib.save(nib.Nifti1Image(array, affine), path)` does; gzip when the name ends in .gz."""
    arr = np.asarray(array.detach().cpu().numpy() if hasattr(array, "detach") else array)
    if arr.dtype == np.float16:
        arr = arr.astype(np.float32)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if arr.dtype not in _CODES:
        raise NotImplementedError(f"dtype {arr.dtype} has no NIfTI-1 datatype code here")
    if not 1 <= arr.ndim <= 7:
        raise ValueError(f"NIfTI-1 holds 1 to 7 dimensions, got {arr.ndim}")
    affine = np.asarray(affine, dtype=np.float64)
    if affine.shape != (4, 4):
        raise ValueError(f"affine must be 4 x 4, got {affine.shape}")
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [arr.ndim] + list(arr.shape) + [1] * (7 - arr.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<2h", hdr, 70, _CODES[arr.dtype], arr.dtype.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(axis=0))
    pixdim = [1.0] + [float(z) for z in zooms] + [1.0] * 4
    struct.pack_into("<8f", hdr, 76, *pixdim)
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)            # vox_offset, scl_slope, scl_inter
    hdr[123] = 2                                                  # xyzt_units: millimetres
    struct.pack_into("<2h", hdr, 252, 0, 2)                       # qform_code unknown, sform_code aligned
    struct.pack_into("<3f", hdr, 268, *[float(v) for v in affine[:3, 3]])
    struct.pack_into("<12f", hdr, 280, *[float(v) for v in affine[:3, :].reshape(-1)])
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\0\0\0\0")                                      # header extension flag: none
        f.write(np.asfortranarray(arr).tobytes(order="F"))


def zscore(volume, mean: float, std: float):
    """(x - mean) / std with the dataset constants of config.yaml:12-15 (data.py:167-171, test_all.py:211-214)."""
    return (volume - mean) / std
