"""Import shim for the read-only reference checkout (test infrastructure only).

`/root/reference/imagen_pytorch3D.py` cannot be imported as shipped: it pulls in
packages that are absent here (einops_exts, kornia, torchmetrics, matplotlib),
a module that is missing from the tree (percept_loss -> MedicalNet) and a module
that fetches from the network at import (imagen_video -> t5).  None of those are
used by the sampling path, so this file plants inert stand-ins in `sys.modules`
 and then imports the real, unmodified reference file.

It is used ONLY by `tests/golden/make_golden.py` and by CPU tests that pin the
oracle against the live reference when `/root/reference` exists.  Nothing here
is reachable from the product package, `bench.py` or the `-m gpu` tests.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_DIR = os.environ.get("DIQT_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "imagen_pytorch3D.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load_reference():
    """Return the reference `imagen_pytorch3D` module (imported once)."""
    if "imagen_pytorch3D" in sys.modules:
        return sys.modules["imagen_pytorch3D"]
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")

    import torch
    from torch import nn
    from einops import rearrange, repeat

    if "einops_exts" not in sys.modules:
        _stub(
            "einops_exts",
            rearrange_many=lambda ts, pattern, **kw: tuple(rearrange(t, pattern, **kw) for t in ts),
            repeat_many=lambda ts, pattern, **kw: tuple(repeat(t, pattern, **kw) for t in ts),
            check_shape=lambda t, pattern, **kw: rearrange(t, f"{pattern} -> {pattern}", **kw),
        )
    if "kornia" not in sys.modules:
        k = _stub("kornia")
        k.augmentation = _stub("kornia.augmentation")
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            tv = _stub("torchvision")
            tv.transforms = _stub("torchvision.transforms")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            m = _stub("matplotlib")
            m.pyplot = _stub("matplotlib.pyplot")
    if "torchmetrics" not in sys.modules:
        try:
            from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity  # noqa: F401
        except Exception:
            class LearnedPerceptualImagePatchSimilarity(nn.Module):  # never constructed (lpips=False)
                def __init__(self, *a, **k):
                    super().__init__()

            tm = _stub("torchmetrics")
            tm.image = _stub("torchmetrics.image")
            tm.image.lpip = _stub(
                "torchmetrics.image.lpip",
                LearnedPerceptualImagePatchSimilarity=LearnedPerceptualImagePatchSimilarity,
            )
    _stub("percept_loss")

    class Unet3D(nn.Module):  # isinstance target only (imagen_pytorch3D.py:1849)
        pass

    _stub("imagen_video", Unet3D=Unet3D, resize_video_to=lambda x, *a, **k: x)

    sys.path.insert(0, REFERENCE_DIR)
    try:
        import imagen_pytorch3D  # type: ignore
    finally:
        sys.path.remove(REFERENCE_DIR)
    # the reference switches autograd anomaly mode on at import (imagen_pytorch3D.py:34)
    torch.autograd.set_detect_anomaly(False)
    return imagen_pytorch3D
