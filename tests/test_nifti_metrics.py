"""On-disk format and metrics either side of the path: NIfTI-1 reader / writer against the standard's field layout,
product metrics against the independent oracle restatement."""
import gzip
import struct

import numpy as np
import pytest
import torch

from diffusioniqt_b200 import metrics as M
from diffusioniqt_b200.nifti import load_nifti, save_nifti, zscore
from diffusioniqt_b200.synth import synthetic_field
from oracle import metrics_oracle as mo


def _affine():
    a = np.array([[0.0, -1.5, 0.0, 90.0], [1.5, 0.0, 0.0, -126.0], [0.0, 0.0, 2.0, -72.0], [0, 0, 0, 1.0]])
    return a


@pytest.mark.parametrize("name", ["vol.nii", "vol.nii.gz"])
@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.uint8, np.float64])
def test_round_trip_and_header_layout(tmp_path, name, dtype):
    rs = np.random.RandomState(0)
    arr = (rs.standard_normal((5, 7, 6)) * 50).astype(dtype)
    path = tmp_path / name
    save_nifti(arr, _affine(), path)
    raw = (gzip.open(path, "rb") if name.endswith(".gz") else open(path, "rb")).read()
    # NIfTI-1 standard (nifti1.h): sizeof_hdr @0 = 348, dim @40, datatype @70, bitpix @72, pixdim @76, vox_offset @108, scl_slope @112,
    # sform_code @254, srow_x @280, magic @344 = "n+1\0"; single-file data starts at vox_offset = 352; voxel data in Fortran order
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\0"
    assert struct.unpack("<8h", raw[40:56]) == (3, 5, 7, 6, 1, 1, 1, 1)
    code, bitpix = struct.unpack("<2h", raw[70:74])
    assert (code, bitpix) == {np.float32: (16, 32), np.int16: (4, 16), np.uint8: (2, 8), np.float64: (64, 64)}[dtype]
    assert struct.unpack("<f", raw[108:112])[0] == 352.0 and struct.unpack("<2h", raw[252:256]) == (0, 2)
    assert np.allclose(struct.unpack("<4f", raw[280:296]), _affine()[0])
    assert np.allclose(struct.unpack("<8f", raw[76:108])[1:4], [1.5, 1.5, 2.0])
    assert len(raw) == 352 + arr.size * arr.itemsize
    first = np.frombuffer(raw, dtype=dtype, count=5, offset=352)
    assert np.array_equal(first, arr[:, 0, 0])                       # x varies fastest
    data, affine, hdr = load_nifti(path)
    assert data.dtype == np.float64 and data.shape == arr.shape and np.array_equal(data, arr.astype(np.float64))
    assert np.allclose(affine, _affine()) and hdr["sform_code"] == 2


def test_reads_big_endian_scaled_and_qform(tmp_path):
    arr = np.arange(24, dtype=">i2").reshape(2, 3, 4, order="F")
    hdr = bytearray(348)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, 2, 3, 4, 1, 1, 1, 1)
    struct.pack_into(">2h", hdr, 70, 4, 16)
    struct.pack_into(">8f", hdr, 76, -1.0, 2.0, 3.0, 4.0, 1, 1, 1, 1)         # qfac -1
    struct.pack_into(">3f", hdr, 108, 352.0, 0.5, 10.0)                       # slope 0.5, inter 10
    struct.pack_into(">2h", hdr, 252, 1, 0)                                   # qform only
    struct.pack_into(">3f", hdr, 256, 0.0, 0.0, 0.0)                          # identity rotation
    struct.pack_into(">3f", hdr, 268, 5.0, 6.0, 7.0)
    hdr[344:348] = b"n+1\0"
    path = tmp_path / "be.nii"
    path.write_bytes(bytes(hdr) + b"\0\0\0\0" + arr.tobytes(order="F"))
    data, affine, h = load_nifti(path)
    assert h["endian"] == ">" and np.array_equal(data, arr.astype(np.float64) * 0.5 + 10.0)
    assert np.allclose(affine, [[2, 0, 0, 5], [0, 3, 0, 6], [0, 0, -4, 7], [0, 0, 0, 1]])


def test_rejects_garbage(tmp_path):
    p = tmp_path / "x.nii"
    p.write_bytes(b"\0" * 400)
    with pytest.raises(ValueError):
        load_nifti(p)
    p.write_bytes(b"short")
    with pytest.raises(ValueError):
        load_nifti(p)
    with pytest.raises(ValueError):
        save_nifti(np.zeros((2, 2, 2), np.float32), np.eye(3), tmp_path / "y.nii")


def test_torch_tensor_and_zscore(tmp_path):
    t = torch.arange(27, dtype=torch.float32).reshape(3, 3, 3)
    save_nifti(t, np.eye(4), tmp_path / "t.nii.gz")
    data, _, _ = load_nifti(tmp_path / "t.nii.gz")
    assert np.array_equal(data, t.numpy().astype(np.float64))
    z = zscore(torch.tensor([271.64814106698583, 648.765]), 271.64814106698583, 377.117173547721)
    assert abs(float(z[0])) < 1e-7


def test_metrics_match_the_oracle_restatement():
    a = synthetic_field((40, 40, 40), 1)
    b = a + 0.1 * synthetic_field((40, 40, 40), 2)
    assert abs(M.psnr(a, b) - mo.psnr(a, b)) < 1e-9
    assert abs(M.ssim3d(a, b) - mo.ssim3d(a, b)) < 1e-9                      # separable vs dense 3-D window
    assert M.ssim3d(a, a) == pytest.approx(1.0)
    with pytest.raises(ValueError):
        M.ms_ssim3d(a, b)                                                    # 40 voxels: the coarsest of the 5 scales would be narrower than the window


def test_ms_ssim_definition():
    """prod_i cs_i^beta_i * ssim_last^beta_last with 2x average pooling between scales, checked scale by scale against a dense-window
    evaluation (3-wide window so that a 48^3 volume has five valid scales)."""
    import torch.nn.functional as F
    K, SIG = 3, 1.5
    a = synthetic_field((48, 48, 48), 3)
    b = a + 0.2 * synthetic_field((48, 48, 48), 4)
    an, bn = (a - a.min()) / (a.max() - a.min()), (b - b.min()) / (b.max() - b.min())
    got = M.ms_ssim3d(an, bn, kernel_size=K, sigma=SIG)
    g = torch.exp(-((torch.arange(K, dtype=torch.float64) - (K - 1) / 2) ** 2) / (2 * SIG ** 2))
    g = g / g.sum()
    k3 = (g[:, None, None] * g[None, :, None] * g[None, None, :])[None, None]
    blur = lambda x: F.conv3d(x, k3)
    p, t, terms = an[None, None].double(), bn[None, None].double(), []
    for i, beta in enumerate(M.MS_SSIM_BETAS):
        mp, mt = blur(p), blur(t)
        cs = (2 * (blur(p * t) - mp * mt) + 0.03 ** 2) / ((blur(p * p) - mp ** 2) + (blur(t * t) - mt ** 2) + 0.03 ** 2)
        ssim = (2 * mp * mt + 0.01 ** 2) / (mp ** 2 + mt ** 2 + 0.01 ** 2) * cs
        assert float(ssim.mean()) == pytest.approx(mo.ssim3d(p[0, 0], t[0, 0], kernel_size=K, sigma=SIG, normalise=False), rel=1e-9)
        terms.append(float(ssim.mean() if i == len(M.MS_SSIM_BETAS) - 1 else cs.mean()) ** beta)
        p, t = F.avg_pool3d(p, 2), F.avg_pool3d(t, 2)
    assert got == pytest.approx(float(np.prod(terms)), rel=1e-9)
    assert 0.0 < got < 1.0 and M.ms_ssim3d(an, an, kernel_size=K) == pytest.approx(1.0)
    with pytest.raises(ValueError):
        M.ms_ssim3d(an, bn)                                                  # default 11-wide window: 48 / 16 = 3 < 11


def test_evaluate_crops_like_the_reference_script():
    """test_all.py:47-85: volumes whose first side is 240 / 256 lose 24 / 32 voxels per face, then (MS-SSIM of the min-max normalised
    volumes, PSNR)."""
    a = synthetic_field((240, 100, 100), 5)
    b = a + 0.1 * synthetic_field((240, 100, 100), 6)
    s, p = M.evaluate(a, b, kernel_size=3)
    ca, cb = a[24:-24, 24:-24, 24:-24], b[24:-24, 24:-24, 24:-24]
    assert p == pytest.approx(mo.psnr(ca, cb), rel=1e-9)
    norm = lambda v: (v - v.min()) / (v.max() - v.min())
    assert s == pytest.approx(M.ms_ssim3d(norm(ca), norm(cb), kernel_size=3), rel=1e-8)
    small = synthetic_field((60, 60, 60), 8)
    assert M.evaluate(small, small, kernel_size=3)[0] == pytest.approx(1.0)      # other sizes: no crop
