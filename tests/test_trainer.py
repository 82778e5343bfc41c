"""ImagenTrainer mirror: checkpoint dictionary of trainer.py:831-860 / 880-945, EMA swap, batch chunking.  CPU only
(sampling itself is covered by the GPU tests; here `Imagen.sample` is replaced by a recorder)."""
import numpy as np
import pytest
import torch

import ref_shim
from cases import MIN_BOUND

KW = dict(dim=32, init_dim=32, dim_mults=(1, 2), num_resnet_blocks=(1, 1), channels=1, lowres_cond=True, init_cross_embed=False,
          attend_at_middle=False, attend_at_enc=(False, False), use_se_attn=True, memory_efficient=False, deep_feature=False,
          boundary=False, batch_sample=False, img_size=8)
IMAGEN_KW = dict(image_sizes=(8, 8), channels=1, timesteps=4, pred_objectives="x_start", dynamic_thresholding=False, min_bound=MIN_BOUND,
                 cond_drop_prob=0.0, p2_loss_weight_gamma=0.0, auto_normalize_img=False)
CONFIGS = {"Data": {"norm": "z-score"}, "Train": {"batch_sample": False}}


def _imagen(seed):
    from diffusioniqt_b200 import Imagen, NullUnet, Unet
    from diffusioniqt_b200.synth import fill_module_
    u = Unet(**KW)
    fill_module_(u, seed=seed)
    return Imagen(unets=(NullUnet(), u), configs=CONFIGS, **IMAGEN_KW)


def test_save_load_round_trip(tmp_path):
    from diffusioniqt_b200.trainer import ImagenTrainer
    a = ImagenTrainer(CONFIGS, imagen=_imagen(1), gradient_accumulation_steps=4, split_valid_from_train=False, lr=3e-4, warmup_steps=10)
    with torch.no_grad():                                  # make the EMA weights differ from the online ones
        for p in a.ema_unets[1].ema_model.parameters():
            p.mul_(0.5)
    a.steps += 7
    path = tmp_path / "model" / "3dimagen.pt"
    a.save(path)
    obj = torch.load(path, weights_only=False)
    assert set(obj) == {"model", "version", "steps", "ema"}
    assert "1.ema_model.init_conv.weight" in obj["ema"] and "1.online_model.init_conv.weight" in obj["ema"] and "0.initted" in obj["ema"]
    b = ImagenTrainer(CONFIGS, imagen=_imagen(2))
    loaded = b.load(path)
    assert loaded["version"] == "1.20.1" and b.num_steps_taken(2) == 7
    for (k, v), (_, w) in zip(a.imagen.state_dict().items(), b.imagen.state_dict().items()):
        assert torch.equal(v, w), k
    for (k, v), (_, w) in zip(a.ema_unets[1].ema_model.state_dict().items(), b.ema_unets[1].ema_model.state_dict().items()):
        assert torch.equal(v, w), k
    assert not torch.equal(b.ema_unets[1].ema_model.init_conv.weight, b.imagen.unets[1].init_conv.weight)
    assert b.load(tmp_path / "missing.pt", noop_if_not_exist=True) is None
    with pytest.raises(AssertionError):
        b.load(tmp_path / "missing.pt")


def test_partial_restore_on_shape_mismatch(tmp_path, capsys):
    from diffusioniqt_b200.trainer import ImagenTrainer
    a = ImagenTrainer(CONFIGS, imagen=_imagen(1), use_ema=False)
    path = tmp_path / "ckpt.pt"
    a.save(path)
    obj = torch.load(path, weights_only=False)
    obj["model"]["unets.1.final_conv.bias"] = torch.zeros(3)          # wrong size: triggers restore_parts (trainer.py:222-233, 899-904)
    torch.save(obj, path)
    b = ImagenTrainer(CONFIGS, imagen=_imagen(2), use_ema=False)
    before = b.imagen.unets[1].final_conv.bias.clone()
    b.load(path)
    assert "Trying partial load" in capsys.readouterr().out
    assert torch.equal(b.imagen.unets[1].final_conv.bias, before)
    assert torch.equal(b.imagen.unets[1].init_conv.weight, a.imagen.unets[1].init_conv.weight)


def test_sample_uses_ema_weights_casts_and_chunks():
    from diffusioniqt_b200.trainer import ImagenTrainer
    t = ImagenTrainer(CONFIGS, imagen=_imagen(1))
    online = t.imagen.unets
    calls = []

    def fake_sample(*args, **kw):
        calls.append((t.imagen.unets[1], kw))
        b = kw["batch_size"]
        return torch.full((b, 1, 8, 8, 8), float(len(calls))), [np.zeros((b, 1))] * 2, [np.ones((b, 1))] * 2

    t.imagen.sample = fake_sample
    lr = np.zeros((5, 1, 8, 8, 8), np.float32)
    img, noisy, x0 = t.sample(batch_size=5, start_image_or_video=lr, start_at_unet_number=2, max_batch_size=2, skip_steps=None,
                              return_all_outputs=False, return_pil_images=False)
    assert [c[1]["batch_size"] for c in calls] == [2, 2, 1]
    assert all(isinstance(c[1]["start_image_or_video"], torch.Tensor) and c[1]["start_image_or_video"].shape[0] == c[1]["batch_size"] for c in calls)
    assert all(c[0] is t.ema_unets[1].ema_model for c in calls) and all(c[1]["use_tqdm"] is False for c in calls)
    assert t.imagen.unets is online                                  # swapped back (trainer.py:999)
    assert img.shape == (5, 1, 8, 8, 8) and img[:, 0, 0, 0, 0].tolist() == [1, 1, 2, 2, 3]
    assert len(noisy) == 2 and noisy[0].shape == (5, 1) and x0[1].shape == (5, 1)
    calls.clear()
    t.sample(batch_size=1, start_image_or_video=torch.zeros(1, 1, 8, 8, 8), start_at_unet_number=2, use_non_ema=True)
    assert calls[0][0] is online[1]
    # training runs on the sm_100a kernels only: on a CPU-resident trainer the forward refuses instead of falling back
    with pytest.raises(RuntimeError, match="CUDA"):
        t(torch.zeros(1, 1, 8, 8, 8), torch.zeros(1, 1, 8, 8, 8), unet_number=2)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present")
def test_loads_a_checkpoint_written_from_the_reference_modules(tmp_path):
    """The dictionary `ImagenTrainer.save` writes (trainer.py:831-860), assembled from the live reference `Imagen` (the reference
    trainer itself needs accelerate / ema_pytorch, absent here): restored by name into this package's modules."""
    from diffusioniqt_b200.synth import fill_module_
    from diffusioniqt_b200.trainer import ImagenTrainer
    ref = ref_shim.load_reference()
    ru = ref.Unet(**KW)
    fill_module_(ru, seed=5)
    rim = ref.Imagen(unets=(ref.NullUnet(), ru), configs=CONFIGS, **IMAGEN_KW)
    ema = {}
    for i, u in enumerate(rim.unets):
        for k, v in u.state_dict().items():
            ema[f"{i}.online_model.{k}"] = v
            ema[f"{i}.ema_model.{k}"] = v * 0.25
        ema[f"{i}.initted"], ema[f"{i}.step"] = torch.tensor([True]), torch.tensor([123])
    path = tmp_path / "3dimagen.pt"
    torch.save(dict(model=rim.state_dict(), version="1.20.1", steps=torch.tensor([0, 9]), ema=ema, optim0={}, scaler0={}), path)
    t = ImagenTrainer(CONFIGS, imagen=_imagen(1))
    t.load(path)                                                       # strict
    mine, theirs = t.imagen.state_dict(), rim.state_dict()
    assert list(mine) == list(theirs)
    assert all(torch.equal(mine[k], theirs[k]) for k in mine)
    assert torch.equal(t.ema_unets[1].ema_model.init_conv.weight, ru.init_conv.weight * 0.25)
    assert int(t.ema_unets[1].step) == 123 and t.num_steps_taken(2) == 9
