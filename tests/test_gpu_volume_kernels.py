"""Patch gather / stitch kernels through `volume.infer_volume` on the GPU, against the loop-by-loop numpy restatement of
data.py:159-162, 192-196 and test_all.py:239-300 (oracle/stitch_oracle.py).  Pure data movement: bit-exact."""
import numpy as np
import pytest
import torch

from diffusioniqt_b200 import volume as V
from oracle import stitch_oracle as so

pytestmark = pytest.mark.gpu


def _volume(shape, seed, holes=True):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(*shape, generator=g)
    if holes:
        v[: shape[0] // 2 + 1] = v.min()                       # a background slab: the patches inside it fall under the 5 % rule
    return v


@pytest.mark.parametrize("shape,P,stride,batch_sample,f,bs", [
    ((40, 40, 40), 16, 8, False, 0, 3),        # plain branch, margins 4, overlapping crops
    ((40, 32, 48), 16, 8, False, 0, 5),        # non-cubic: the reference uses shape[-1] for every axis
    ((48, 48, 48), 16, 16, False, 0, 4),       # stride == patch: no cropping
    ((36, 36, 36), 12, 8, True, 3, 1),         # batch_sample branch, patches travel as 27 sub-volumes
    ((32, 32, 32), 16, 8, True, 2, 2),
    ((64, 64, 64), 32, 16, False, 0, 7),
])
def test_gather_and_stitch_match_the_reference_loops(shape, P, stride, batch_sample, f, bs):
    low = _volume(shape, 1)
    raw = low - low.min()
    calls = []

    def fake(lr):                                  # a stand-in sampler that is easy to restate on the host
        calls.append(tuple(lr.shape))
        return lr * 2 + 1

    res = V.infer_volume(fake, low.cuda(), patch=P, overlap=stride, raw_lowres=raw.cuda(), batch_size=bs, fill_value=-3.0,
                         batch_sample=batch_sample, sub_f=f)
    idxs = [i for i in so.patch_index_list(shape, P, stride) if not so.is_skipped(raw.numpy(), i, P)]
    outs = [(low[i:i + P, j:j + P, k:k + P] * 2 + 1).numpy() for i, j, k in idxs]
    want = so.background_mask(so.stitch(np.full(shape, -3.0, np.float32), outs, idxs, P, stride, batch_sample), low.numpy())
    assert res.n_patches == len(idxs) and res.n_skipped > 0
    assert torch.equal(res.volume.cpu(), torch.from_numpy(want))
    side = P // f if f > 1 else P
    assert all(c[1:] == (1, side, side, side) for c in calls)
    assert sum(c[0] for c in calls) == len(idxs) * (f ** 3 if f > 1 else 1)


def test_gathered_sub_volumes_are_the_reference_order():
    """diqt_gather_patches with sub_f: the same 27 sub-volumes as convertVolume2subVolume (pinned oracle helper)."""
    from diffusioniqt_b200 import lib as L
    from oracle import unet_oracle as uo
    lib = L.load()
    vol = _volume((30, 28, 26), 2, holes=False).cuda()
    origins = torch.tensor([[0, 0, 0], [6, 4, 2], [3, 1, 0]], dtype=torch.int32, device="cuda")
    P, f = 24, 3
    out = torch.empty(3 * 27, 1, 8, 8, 8, device="cuda")
    L.check(lib.diqt_gather_patches(vol.data_ptr(), 30, 28, 26, origins.data_ptr(), 3, P, f, out.data_ptr(), L.current_stream()), "gather")
    for b, (i, j, k) in enumerate(origins.tolist()):
        want = uo.split_sub_volumes(vol[i:i + P, j:j + P, k:k + P].cpu()[None, None], f)
        assert torch.equal(out[b * 27:(b + 1) * 27].cpu(), want)
