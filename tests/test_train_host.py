"""Host-side logic of the training step (diffusioniqt_b200/train.py) against PyTorch autograd on the CPU: the GroupNorm / FiLM / Mish
coefficient algebra between the reduce and the apply kernel (the two kernels are emulated in torch here), the channels-last pixel
(un)shuffles, the flipped-weight data gradient and the EMA schedule."""
import torch
import torch.nn.functional as F

from diffusioniqt_b200.train import _flip_t, _shuffle_cl, _unshuffle_cl, gn_backward_coefficients
from oracle.unet_oracle import pixel_shuffle3d, pixel_unshuffle3d


def _mish_grad(w):
    sp = F.softplus(w)
    t = torch.tanh(sp)
    return t + w * torch.sigmoid(w) * (1 - t * t)


def test_groupnorm_film_mish_backward_algebra_matches_autograd():
    torch.manual_seed(0)
    n, c, G, S = 2, 16, 4, 6
    x = (torch.randn(n, c, S, S, S, dtype=torch.float64) * 1.7 + 0.4).requires_grad_(True)
    gamma = torch.randn(c, dtype=torch.float64).requires_grad_(True)
    beta = torch.randn(c, dtype=torch.float64).requires_grad_(True)
    film = (torch.randn(n, 2 * c, dtype=torch.float64) * 0.5).requires_grad_(True)
    dy = torch.randn(n, c, S, S, S, dtype=torch.float64)
    scale, shift = film[:, :c, None, None, None], film[:, c:, None, None, None]
    y = F.mish(F.group_norm(x, G, gamma, beta, eps=1e-5) * (scale + 1) + shift)
    y.backward(dy)
    # what the forward kernels leave behind: group mean / rstd and the folded affine w = a x + b
    xd = x.detach()
    xg = xd.reshape(n, G, -1)
    mean, var = xg.mean(dim=2), xg.var(dim=2, unbiased=False)
    rstd = 1 / torch.sqrt(var + 1e-5)
    cpg = c // G
    mu_c, r_c = mean.repeat_interleave(cpg, 1), rstd.repeat_interleave(cpg, 1)
    k = 1 + film.detach()[:, :c]
    a = r_c * gamma.detach() * k
    b = (beta.detach() - mu_c * r_c * gamma.detach()) * k + film.detach()[:, c:]
    # diqt_bwd_reduce (mode 1)
    w = a[:, :, None, None, None] * xd + b[:, :, None, None, None]
    dw = dy * _mish_grad(w)
    S1, S2x = dw.sum(dim=(2, 3, 4)), (dw * xd).sum(dim=(2, 3, 4))
    c1, c2, c3, dgamma, dbeta, dfilm = gn_backward_coefficients(S1, S2x, mean, rstd, gamma.detach(), beta.detach(), film.detach(), S ** 3, G)
    # diqt_bwd_apply (mode 1)
    dx = c1.double()[:, :, None, None, None] * dw + c2.double()[:, :, None, None, None] * xd + c3.double()[:, :, None, None, None]
    assert torch.allclose(dx, x.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(dgamma, gamma.grad, rtol=1e-9, atol=1e-9)
    assert torch.allclose(dbeta, beta.grad, rtol=1e-9, atol=1e-9)
    assert torch.allclose(dfilm.double(), film.grad, rtol=1e-5, atol=1e-5)
    # without FiLM
    x2 = xd.clone().requires_grad_(True)
    y2 = F.mish(F.group_norm(x2, G, gamma.detach(), beta.detach(), eps=1e-5))
    y2.backward(dy)
    a2, b2 = r_c * gamma.detach(), beta.detach() - mu_c * r_c * gamma.detach()
    dw2 = dy * _mish_grad(a2[:, :, None, None, None] * xd + b2[:, :, None, None, None])
    c1, c2, c3, _, _, none = gn_backward_coefficients(dw2.sum(dim=(2, 3, 4)), (dw2 * xd).sum(dim=(2, 3, 4)), mean, rstd, gamma.detach(), beta.detach(),
                                                     None, S ** 3, G)
    dx2 = c1.double()[:, :, None, None, None] * dw2 + c2.double()[:, :, None, None, None] * xd + c3.double()[:, :, None, None, None]
    assert none is None and torch.allclose(dx2, x2.grad, rtol=1e-4, atol=1e-6)


def test_channels_last_pixel_shuffles_match_the_oracle():
    torch.manual_seed(1)
    x = torch.randn(2, 3, 4, 6, 8)
    cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()
    assert torch.equal(_unshuffle_cl(cl(x)), cl(pixel_unshuffle3d(x)))
    y = torch.randn(2, 16, 2, 3, 4)
    assert torch.equal(_shuffle_cl(cl(y)), cl(pixel_shuffle3d(y)))
    assert torch.equal(_shuffle_cl(_unshuffle_cl(cl(x))), cl(x))


def test_data_gradient_is_a_convolution_with_flipped_transposed_weights():
    torch.manual_seed(2)
    x = torch.randn(1, 5, 6, 7, 8, dtype=torch.float64, requires_grad=True)
    w = torch.randn(4, 5, 3, 3, 3, dtype=torch.float64)
    dy = torch.randn(1, 4, 6, 7, 8, dtype=torch.float64)
    F.conv3d(x, w, None, padding=1).backward(dy)
    assert torch.allclose(F.conv3d(dy, _flip_t(w), None, padding=1), x.grad, atol=1e-10)
    w1 = torch.randn(4, 5, 1, 1, 1, dtype=torch.float64)
    x.grad = None
    F.conv3d(x, w1).backward(dy)
    assert torch.allclose(F.conv3d(dy, _flip_t(w1)), x.grad, atol=1e-10)


def test_ema_schedule_of_the_trainer():
    from diffusioniqt_b200.trainer import ImagenTrainer
    d = ImagenTrainer._ema_decay
    class T: EMA_BETA, EMA_UPDATE_AFTER, EMA_UPDATE_EVERY, EMA_INV_GAMMA, EMA_POWER = 0.9999, 100, 10, 1.0, 2.0 / 3.0
    assert d(T, 100) == 0. and d(T, 101) == 0.
    assert abs(d(T, 110) - (1 - (1 + 9) ** (-2 / 3))) < 1e-12
    assert d(T, 10 ** 9) == 0.9999


def _allreduce_worker(rank, world, port, ret):
    import os
    import torch.distributed as dist
    from diffusioniqt_b200.train import allreduce_gradients
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((3, 5), float(rank + 1))
    params[1].grad = torch.arange(7.) * (rank + 1)
    if rank == 0:
        params[2].grad = torch.ones(2, 2)          # rank 1 has no gradient for this one
    allreduce_gradients(params)
    ret[rank] = [p.grad.clone() for p in params]
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """ImagenTrainer.update averages the gradients over the ranks with one flat all-reduce (NCCL on the GPU box, gloo here)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_allreduce_worker, args=(2, port, ret), nprocs=2, join=True)
    for r in (0, 1):
        g = ret[r]
        assert torch.equal(g[0], torch.full((3, 5), 1.5)) and torch.equal(g[1], torch.arange(7.) * 1.5) and torch.equal(g[2], torch.full((2, 2), 0.5))


def test_learning_rate_schedule_matches_torch_cosine_annealing_and_linear_warmup():
    """ImagenTrainer.scheduled_lr against torch.optim.lr_scheduler.CosineAnnealingLR (trainer.py:368-369) and the linear warm-up factor."""
    from diffusioniqt_b200.trainer import ImagenTrainer
    class T:
        _optim_args = dict(lr=(1e-4, 3e-4), warmup_steps=(None, 50), cosine_decay_max_steps=(None, 200))
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=3e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=200, eta_min=3e-4 * 0.001)
    for step in range(260):
        want = opt.param_groups[0]["lr"] * min(1.0, (step + 1) / 50)
        assert abs(ImagenTrainer.scheduled_lr(T, 1, step) - want) < 1e-12 * max(1, step) + 1e-15, step
        opt.step(); sched.step()
    assert ImagenTrainer.scheduled_lr(T, 0, 123) == 1e-4


def test_adam_state_round_trips_through_the_torch_optimizer_format():
    """AdamState.state_dict() is torch.optim.Adam's layout (what the reference stores under 'optim{i}', trainer.py:858): a torch optimizer
    loads it, and AdamState loads what a torch optimizer wrote."""
    from diffusioniqt_b200.train import AdamState
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ours = AdamState(ps, lr=2e-4, betas=(0.9, 0.99), eps=1e-8)
    ours.steps = 7
    for m, v in zip(ours.m, ours.v):
        m.copy_(torch.randn_like(m)); v.copy_(torch.rand_like(v))
    sd = ours.state_dict()
    ref = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0)
    ref.load_state_dict(sd)
    got = ref.state_dict()
    assert got["param_groups"][0]["lr"] == 2e-4 and tuple(got["param_groups"][0]["betas"]) == (0.9, 0.99)
    for i in range(2):
        assert torch.equal(got["state"][i]["exp_avg"], ours.m[i]) and torch.equal(got["state"][i]["exp_avg_sq"], ours.v[i])
        assert float(got["state"][i]["step"]) == 7.0
    back = AdamState(ps, lr=1.0)
    back.load_state_dict(got)
    assert back.steps == 7 and back.lr == 2e-4 and all(torch.equal(a, b) for a, b in zip(back.m, ours.m))
    fresh = AdamState(ps)
    assert fresh.state_dict()["state"] == {}
