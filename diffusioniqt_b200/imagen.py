"""Host-side mirror of the reference diffusion wrapper: `GaussianDiffusionContinuousTimes`
(/root/reference/imagen_pytorch3D.py:236-357) and the sampling half of `Imagen` (:1741-2274).

`Imagen.sample(...)` keeps the reference's signature and 3-tuple return.  One sampler step is
    torch.randn (noise, same RNG call order as the reference)  ->  U-Net engine (libdiqt_b200 kernels)
    ->  final conv fused with clamp + posterior mean + noise  ->  step counter += 1
captured once in a CUDA graph and replayed `timesteps` times.  Schedule scalars are evaluated
with the same torch expressions the reference uses, once per sampler, into a device table.

Deviations, all documented in DESIGN.md: the per-step `.cpu().numpy()` trajectory copies
(:2147-2153) are opt-in (`keep_trajectory=True`); training (`forward`, `p_losses`), text
conditioning and inpainting are out of scope and raise.
"""
from __future__ import annotations

import math
import os
from contextlib import contextmanager, nullcontext
from typing import Optional

import torch
from torch import nn

from . import lib as L
from .unet import NullUnet, Unet, _cast_tuple


def _log(t, eps=1e-20):
    return torch.log(t.clamp(min=eps))


def alpha_cosine_log_snr(t, s: float = 0.008):
    # imagen_pytorch3D.py:229-231
    return -_log((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1, eps=1e-5)


def beta_linear_log_snr(t):
    # imagen_pytorch3D.py:225-227
    return -torch.log(torch.special.expm1(1e-4 + 10 * (t ** 2)))


def log_snr_to_alpha_sigma(log_snr):
    # imagen_pytorch3D.py:233-234
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


def _pad_dims_to(x, t):
    return t.reshape(*t.shape, *((1,) * (x.dim() - t.dim())))


class GaussianDiffusionContinuousTimes(nn.Module):
    """Noise schedule helper with the reference's method names (sampling subset + q_sample)."""

    def __init__(self, *, noise_schedule, timesteps=1000):
        super().__init__()
        if noise_schedule == "linear":
            self.log_snr = beta_linear_log_snr
        elif noise_schedule == "cosine":
            self.log_snr = alpha_cosine_log_snr
        else:
            raise ValueError(f'invalid noise schedule {noise_schedule}')
        self.num_timesteps = timesteps

    def get_times(self, batch_size, noise_level, *, device):
        return torch.full((batch_size,), noise_level, device=device, dtype=torch.float32)

    def sample_random_times(self, batch_size, *, device):
        return torch.zeros((batch_size,)).float().uniform_(0, 1).to(device)

    def get_condition(self, times):
        return None if times is None else self.log_snr(times)

    def get_sampling_timesteps(self, batch, *, device):
        # imagen_pytorch3D.py:261-266: tuple of (2, batch) tensors, one per step
        times = torch.linspace(1., 0., self.num_timesteps + 1, device=device)
        times = times[None, :].expand(batch, -1)
        times = torch.stack((times[:, :-1], times[:, 1:]), dim=0)
        return times.unbind(dim=-1)

    def q_posterior(self, x_start, x_t, t, *, t_next=None):
        # imagen_pytorch3D.py:290-309 (schedule-level math; the fused kernel applies it per voxel)
        if t_next is None:
            t_next = (t - 1. / self.num_timesteps).clamp(min=0.)
        log_snr, log_snr_next = (_pad_dims_to(x_t, v) for v in (self.log_snr(t), self.log_snr(t_next)))
        alpha, sigma = log_snr_to_alpha_sigma(log_snr)
        alpha_next, sigma_next = log_snr_to_alpha_sigma(log_snr_next)
        c = -torch.special.expm1(log_snr - log_snr_next)
        posterior_mean = alpha_next * (x_t * (1 - c) / alpha + c * x_start)
        posterior_variance = (sigma_next ** 2) * c
        return posterior_mean, posterior_variance, _log(posterior_variance, eps=1e-20)

    def q_sample(self, x_start, t, noise=None):
        # imagen_pytorch3D.py:311-323
        dtype = x_start.dtype
        if isinstance(t, float):
            t = torch.full((x_start.shape[0],), t, device=x_start.device, dtype=dtype)
        noise = noise if noise is not None else torch.randn_like(x_start)
        log_snr = self.log_snr(t).type(dtype)
        alpha, sigma = log_snr_to_alpha_sigma(_pad_dims_to(x_start, log_snr))
        return alpha * x_start + sigma * noise, log_snr, alpha, sigma


_OBJECTIVES = {"x_start": 0, "noise": 1, "v": 2}


_DEBUG_SYNC = os.environ.get("DIQT_DEBUG_SYNC_REPLAY", "")   # diagnostics only: host sync before / after every graph replay


class _SamplerState:
    """Device tables + captured graph of one (engine, schedule) pair."""

    def __init__(self):
        self.graph = None
        self.key = None


class Imagen(nn.Module):
    """Drop-in for `imagen_pytorch3D.Imagen` on the sampling path (constructor keywords :1742-1768)."""

    def __init__(
        self,
        unets,
        configs,
        *,
        image_sizes,
        min_bound=0,
        channels=3,
        timesteps=1000,
        cond_drop_prob=0.1,
        loss_type='l2',
        noise_schedules='cosine',
        pred_objectives='noise',
        lowres_noise_schedule='linear',
        lowres_sample_noise_level=0.2,
        per_sample_random_aug_noise_level=False,
        auto_normalize_img=False,
        p2_loss_weight_gamma=0.5,
        p2_loss_weight_k=1,
        dynamic_thresholding=True,
        dynamic_thresholding_percentile=0.95,
        only_train_unet_number=None,
        temporal_downsample_factor=1,
        lpips=False,
        medlpips=False,
        boundary=False,
    ):
        super().__init__()
        self.configs = configs
        self.boundary = boundary
        self.use_lpips = bool(lpips or medlpips)   # the perceptual term (:2366-2381) needs a pretrained VGG: sampling ignores it, training refuses
        if loss_type not in ('l1', 'l2', 'huber'):
            raise NotImplementedError()
        self.loss_type = loss_type
        self.min_bound = min_bound
        self.condition_on_text = False
        self.unconditional = True
        self.channels = channels

        unets = _cast_tuple(unets)
        num_unets = len(unets)
        timesteps = _cast_tuple(timesteps, num_unets)

        noise_schedules = _cast_tuple(noise_schedules)
        noise_schedules = (*noise_schedules, *(('cosine',) * max(0, 2 - len(noise_schedules))))
        noise_schedules = (*noise_schedules, *(('linear',) * max(0, num_unets - len(noise_schedules))))
        self.noise_schedulers = nn.ModuleList([
            GaussianDiffusionContinuousTimes(noise_schedule=ns, timesteps=ts) for ts, ns in zip(timesteps, noise_schedules)])
        self.lowres_noise_schedule = GaussianDiffusionContinuousTimes(noise_schedule=lowres_noise_schedule)
        self.pred_objectives = _cast_tuple(pred_objectives, num_unets)

        self.unets = nn.ModuleList([])
        self.unet_being_trained_index = -1
        self.only_train_unet_number = only_train_unet_number
        for ind, one_unet in enumerate(unets):
            assert isinstance(one_unet, (Unet, NullUnet))
            one_unet = one_unet.cast_model_parameters(lowres_cond=not ind == 0, channels=self.channels, channels_out=self.channels)
            self.unets.append(one_unet)

        image_sizes = _cast_tuple(image_sizes)
        self.image_sizes = image_sizes
        assert num_unets == len(image_sizes), f'you did not supply the correct number of u-nets ({len(unets)}) for resolutions {image_sizes}'
        self.sample_channels = _cast_tuple(self.channels, num_unets)

        temporal_downsample_factor = _cast_tuple(temporal_downsample_factor, num_unets)
        self.temporal_downsample_factor = temporal_downsample_factor
        assert temporal_downsample_factor[-1] == 1, 'downsample factor of last stage must be 1'

        lowres_conditions = tuple(map(lambda t: t.lowres_cond, self.unets))
        assert lowres_conditions == (False, *((True,) * (num_unets - 1))), \
            'the first unet must be unconditioned (by low resolution image), and the rest of the unets must have `lowres_cond` set to True'

        self.lowres_sample_noise_level = lowres_sample_noise_level
        self.per_sample_random_aug_noise_level = per_sample_random_aug_noise_level
        self.cond_drop_prob = cond_drop_prob
        self.can_classifier_guidance = cond_drop_prob > 0.
        if auto_normalize_img:
            self.normalize_img = lambda img: img * 2 - 1
            self.unnormalize_img = lambda img: (img + 1) * 0.5
        else:
            self.normalize_img = self.unnormalize_img = lambda img: img
        self.dynamic_thresholding = _cast_tuple(dynamic_thresholding, num_unets)
        self.dynamic_thresholding_percentile = dynamic_thresholding_percentile
        self.p2_loss_weight_k = p2_loss_weight_k
        self.p2_loss_weight_gamma = _cast_tuple(p2_loss_weight_gamma, num_unets)
        self.register_buffer('_temp', torch.tensor([0.]), persistent=False)

        # ---- execution options of this implementation (not in the reference)
        self.keep_trajectory = False   # True: append img / x_start to host lists every step like :2147-2153
        self.use_cuda_graph = os.environ.get("DIQT_DISABLE_CUDA_GRAPH", "0") != "1"   # profiling runs set the variable
        self.return_host_lists = True  # False: skip the two final host copies too (device-resident benchmarking)
        self.noise_override = None     # tests: an iterable of tensors consumed instead of torch.randn (draw order of the reference)
        self._samplers = {}
        self.to(next(self.unets.parameters()).device)

    # ------------------------------------------------------------------ device bookkeeping (API parity)
    @property
    def device(self):
        return self._temp.device

    def get_unet(self, unet_number):
        assert 0 < unet_number <= len(self.unets)
        return self.unets[unet_number - 1]

    def reset_unets_all_one_device(self, device=None):
        # the reference shuffles unets between CPU and GPU on every call (:1941-1962); weights stay resident here
        device = device if device is not None else self.device
        self.unets.to(device)
        self.unet_being_trained_index = -1

    @contextmanager
    def one_unet_in_gpu(self, unet_number=None, unet=None):
        yield

    def state_dict(self, *args, **kwargs):
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._samplers = {}
        for u in self.unets:
            if isinstance(u, Unet):
                u.invalidate_engines()
        return super().load_state_dict(*args, **kwargs)

    def _clamp_bounds(self):
        if self.configs['Data']['norm'] == 'min-max':
            return -1.0, 1.0
        return float(self.min_bound), float('inf')

    # ------------------------------------------------------------------ schedule table
    def _schedule_pairs(self, noise_scheduler, skip_steps, device):
        """(t, t_next) of every step as two 1-D tensors, after the skip rule of :2103-2107."""
        times = torch.linspace(1., 0., noise_scheduler.num_timesteps + 1, device=device)
        t, t_next = times[:-1], times[1:]
        skip = skip_steps or 0
        if skip > 1:
            idx = list(range(0, t.shape[0], skip)) + [t.shape[0] - 1]
            idx = torch.tensor(idx, device=device)
            t, t_next = t[idx], t_next[idx]
        return t, t_next

    def _build_schedule(self, noise_scheduler, skip_steps, pred_objective, device):
        t, t_next = self._schedule_pairs(noise_scheduler, skip_steps, device)
        log_snr, log_snr_next = noise_scheduler.log_snr(t), noise_scheduler.log_snr(t_next)
        alpha, sigma = log_snr_to_alpha_sigma(log_snr)
        alpha_next, sigma_next = log_snr_to_alpha_sigma(log_snr_next)
        c = -torch.special.expm1(log_snr - log_snr_next)
        log_var = _log((sigma_next ** 2) * c, eps=1e-20)
        nonzero = 1 - (t_next == 0).float()
        noise_scale = nonzero * (0.5 * log_var).exp()
        lo, hi = self._clamp_bounds()
        table = torch.stack([alpha, sigma, c, alpha_next, noise_scale, torch.full_like(alpha, lo), torch.full_like(alpha, hi),
                             torch.full_like(alpha, float(_OBJECTIVES[pred_objective]))], dim=1).contiguous().float()
        return table, log_snr.float().contiguous()

    # ------------------------------------------------------------------ reference-shaped primitives
    @torch.no_grad()
    def p_mean_variance(self, unet, x, t, *, noise_scheduler, cond_images=None, lowres_cond_img=None, self_cond=None, cond_scale=1.,
                        model_output=None, t_next=None, pred_objective='noise', dynamic_threshold=True):
        """Same contract as :1976-2030: ((mean, variance, log_variance), x_start)."""
        assert not (cond_scale != 1. and not self.can_classifier_guidance), \
            'imagen was not trained with conditional dropout, and thus one cannot use classifier free guidance (cond_scale anything other than 1)'
        pred = model_output if model_output is not None else unet.forward_with_cond_scale(
            x, t, noise_scheduler.get_condition(t), cond_images=cond_images, cond_scale=cond_scale, lowres_cond_img=lowres_cond_img, self_cond=self_cond)
        x_start = self._x_start_from_pred(pred, x, t, noise_scheduler, pred_objective, dynamic_threshold)
        return noise_scheduler.q_posterior(x_start=x_start, x_t=x, t=t, t_next=t_next), x_start

    def _x_start_from_pred(self, pred, x, t, noise_scheduler, pred_objective, dynamic_threshold):
        if pred_objective == 'noise':
            alpha, sigma = log_snr_to_alpha_sigma(_pad_dims_to(x, noise_scheduler.log_snr(t)))
            x_start = (x - sigma * pred) / alpha.clamp(min=1e-8)
        elif pred_objective == 'x_start':
            x_start = pred
        elif pred_objective == 'v':
            alpha, sigma = log_snr_to_alpha_sigma(_pad_dims_to(x, noise_scheduler.log_snr(t)))
            x_start = alpha * x - sigma * pred
        else:
            raise ValueError(f'unknown objective {pred_objective}')
        if dynamic_threshold:
            s = torch.quantile(x_start.reshape(x_start.shape[0], -1).abs(), self.dynamic_thresholding_percentile, dim=-1)
            s = s.clamp(min=1.) if self.configs['Data']['norm'] == 'min-max' else s.clamp(min=self.min_bound)
            s = _pad_dims_to(x_start, s)
            return x_start.clamp(-s, s) / s
        lo, hi = self._clamp_bounds()
        return x_start.clamp(min=lo, max=hi if math.isfinite(hi) else None)

    @torch.no_grad()
    def p_sample(self, unet, x, t, *, noise_scheduler, t_next=None, cond_images=None, cond_scale=1., self_cond=None, lowres_cond_img=None,
                 pred_objective='noise', dynamic_threshold=True):
        """One ancestral step with the reference's contract (:2032-2056): returns (x_next, x_start)."""
        b = x.shape[0]
        (model_mean, _, model_log_variance), x_start = self.p_mean_variance(
            unet, x=x, t=t, t_next=t_next, noise_scheduler=noise_scheduler, cond_images=cond_images, cond_scale=cond_scale,
            lowres_cond_img=lowres_cond_img, self_cond=self_cond, pred_objective=pred_objective, dynamic_threshold=dynamic_threshold)
        noise = torch.randn_like(x)
        nonzero_mask = (1 - (t_next == 0).float()).reshape(b, *((1,) * (x.dim() - 1)))
        return model_mean + nonzero_mask * (0.5 * model_log_variance).exp() * noise, x_start

    # ------------------------------------------------------------------ the hot loop
    @torch.no_grad()
    def p_sample_loop(self, unet, shape, *, noise_scheduler, lowres_cond_img=None, cond_images=None, inpaint_images=None, inpaint_masks=None,
                      inpaint_resample_times=5, init_images=None, skip_steps=None, cond_scale=1, pred_objective='noise',
                      dynamic_threshold=True, use_tqdm=True):
        if inpaint_images is not None or inpaint_masks is not None:
            raise NotImplementedError("inpainting is broken in the reference (undefined right_pad_dims_to_datatype, :2143) and not built here")
        # Classifier-free guidance (:1540-1552, :1993): the reference runs a second forward with cond_drop_prob = 1, but its 3-D Unet.forward
        # never reads cond_drop_prob (there is no text conditioning to drop), so null_logits == logits bit for bit and
        # null + (logits - null) * cond_scale == logits.  The guarded assert is kept, the redundant forward is not run.
        assert not (cond_scale != 1. and not self.can_classifier_guidance), \
            'imagen was not trained with conditional dropout, and thus one cannot use classifier free guidance (cond_scale anything other than 1)'
        device = self.device
        if device.type != 'cuda':
            raise RuntimeError("Imagen.sample runs only on a CUDA device (sm_100a kernels; there is no CPU fallback)")
        lib = L.load()
        batch = shape[0]
        eng = unet.engine_for(batch, shape[2:], device)

        # cached on the engine (never keyed by id(): a rebuilt engine can reuse the address of a dead one while its buffers moved)
        key = (noise_scheduler.log_snr.__name__, noise_scheduler.num_timesteps, skip_steps or 0, pred_objective, bool(dynamic_threshold),
               self._clamp_bounds())
        st = eng.sampler_cache.get(key)
        if st is None:
            st = _SamplerState()
            st.table, st.log_snr = self._build_schedule(noise_scheduler, skip_steps, pred_objective, device)
            st.t = self._schedule_pairs(noise_scheduler, skip_steps, device)[0]
            st.table_plain = st.table.clone()
            st.table_plain[:, 5], st.table_plain[:, 6], st.table_plain[:, 7] = float('-inf'), float('inf'), 0.0
            st.steps = st.table.shape[0]
            st.step = torch.zeros(1, dtype=torch.int32, device=device)
            st.noise = torch.empty(shape, dtype=torch.float32, device=device)
            st.x0 = torch.empty(shape, dtype=torch.float32, device=device)
            eng.sampler_cache[key] = st
        nsteps = st.steps

        inj = iter(self.noise_override) if self.noise_override is not None else None

        def draw(dst=None):
            """The reference's RNG calls, in its order (:2080, :2051); tests may inject a recorded sequence instead."""
            if inj is not None:
                v = next(inj).to(device=device, dtype=torch.float32)
                return v.clone() if dst is None else dst.copy_(v)
            return torch.randn(shape, device=device) if dst is None else dst.normal_()

        img = draw()                                                             # :2080 (RNG draw #0)
        if init_images is not None:
            img += init_images                                                   # :2084-2085
        eng.load_inputs(img, lowres_cond_img, cond_images)
        eng.set_condition(st.log_snr)                                            # time MLPs for every step at once
        if getattr(st, "film_gen", None) != eng.film_gen:                        # the FiLM table moved: the captured graph reads the old one
            st.graph, st.film_gen = None, eng.film_gen
        eng.film_row_ptr, eng.film_stride_n = st.step.data_ptr(), 0
        st.step.zero_()
        x_t = eng.x_in                                                           # sampler state lives in the engine's input buffer
        count = x_t.numel()

        def one_step(i=None):
            eng.run_body()
            if dynamic_threshold:
                # torch.quantile is a sort: kept as a torch call outside the fused kernel
                eng.run_final(fused=False)
                t_i = st.t[i].expand(batch)
                st.x0.copy_(self._x_start_from_pred(eng.pred, x_t, t_i, noise_scheduler, pred_objective, True))
                # the thresholded x_start goes through the posterior update as an 'x_start' prediction with no clamp
                L.check(lib.diqt_ddpm_update(st.x0.data_ptr(), st.table_plain.data_ptr(), st.step.data_ptr(), x_t.data_ptr(),
                                             st.noise.data_ptr(), x_t.data_ptr(), 0, count, L.current_stream()), "ddpm_update")
            else:
                eng.run_final(fused=True, sched=st.table.data_ptr(), step=st.step.data_ptr(), noise=st.noise.data_ptr(), x0=st.x0.data_ptr())
            L.check(lib.diqt_advance_step(st.step.data_ptr(), L.current_stream()), "advance_step")

        traj_x, traj_x0 = [], []
        use_graph = self.use_cuda_graph and not dynamic_threshold
        if use_graph and st.graph is None:
            # warm-up on a side stream (lazy module loading, cudaFuncSetAttribute), then capture one step
            saved = x_t.clone()
            st.noise.zero_()
            s = torch.cuda.Stream(device=device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                one_step(0)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize(device)
            x_t.copy_(saved)
            st.step.zero_()
            g = torch.cuda.CUDAGraph()
            before = L.launch_count()
            with torch.cuda.graph(g):                                            # capture records, it does not execute
                one_step(0)
            st.graph_launches = L.launch_count() - before                        # kernels of ours inside one replay
            st.graph = g
        for i in range(nsteps):
            draw(st.noise)                                                       # == torch.randn_like(x) of :2051, every step
            if use_graph:
                if _DEBUG_SYNC in ("before", "both"):
                    torch.cuda.synchronize(device)
                st.graph.replay()
                if _DEBUG_SYNC in ("after", "both"):
                    torch.cuda.synchronize(device)
            else:
                one_step(i)
            if self.keep_trajectory:
                traj_x.append(x_t.cpu().numpy())                                 # :2148-2149
                traj_x0.append(st.x0.cpu().numpy())
        if self.return_host_lists:
            traj_x.append(x_t.cpu().numpy())                                     # :2151-2152
            traj_x0.append(st.x0.cpu().numpy())
        self.last_graph_launches = getattr(st, "graph_launches", 0)
        lo, hi = self._clamp_bounds()
        img = x_t.clone()
        L.check(lib.diqt_clamp(img.data_ptr(), img.numel(), lo, hi, L.current_stream()), "clamp")                # :2154-2157
        return self.unnormalize_img(img), traj_x, traj_x0

    @torch.no_grad()
    def sample(self, text_masks=None, text_embeds=None, video_frames=None, cond_images=None, inpaint_images=None, inpaint_masks=None,
               inpaint_resample_times=5, init_images=None, skip_steps=None, batch_size=1, cond_scale=1., lowres_sample_noise_level=None,
               start_at_unet_number=1, start_image_or_video=None, stop_at_unet_number=None, return_all_outputs=False,
               return_pil_images=False, device=None, use_tqdm=True):
        """Same signature and `(img, list_noisy, list_x0)` return as the reference (:2165-2274)."""
        was_training = self.training
        self.eval()
        try:
            device = device if device is not None else self.device
            self.reset_unets_all_one_device(device=device)
            if cond_images is not None and cond_images.dtype == torch.uint8:
                cond_images = cond_images / 255
            assert text_embeds is None and text_masks is None, 'this model is unconditional (no text conditioning)'
            outputs = []
            num_unets = len(self.unets)
            cond_scale = _cast_tuple(cond_scale, num_unets)
            init_images = list(_cast_tuple(init_images, num_unets))
            skip_steps = _cast_tuple(skip_steps, num_unets)
            img = None
            if start_at_unet_number > 1:
                assert start_at_unet_number <= num_unets, 'must start a unet that is less than the total number of unets'
                assert stop_at_unet_number is None or start_at_unet_number <= stop_at_unet_number
                assert start_image_or_video is not None, 'starting image or video must be supplied if only doing upscaling'
                img = start_image_or_video
            lst_pred_noisy, lst_pred = [], []
            for unet_number, unet, channel, image_size, noise_scheduler, pred_objective, dynamic_threshold, unet_cond_scale, unet_init_images, unet_skip_steps in zip(
                    range(1, num_unets + 1), self.unets, self.sample_channels, self.image_sizes, self.noise_schedulers, self.pred_objectives,
                    self.dynamic_thresholding, cond_scale, init_images, skip_steps):
                if unet_number < start_at_unet_number:
                    continue
                assert not isinstance(unet, NullUnet), 'one cannot sample from null / placeholder unets'
                lowres_cond_img = None
                if unet.lowres_cond:
                    assert img is not None, 'low resolution conditioning image must be present'
                    lowres_cond_img = img.to(device=device, dtype=torch.float32)
                shape = (batch_size, self.channels, image_size, image_size, image_size)
                img, lst_pred_noisy, lst_pred = self.p_sample_loop(
                    unet, shape, cond_images=cond_images, inpaint_images=inpaint_images, inpaint_masks=inpaint_masks,
                    inpaint_resample_times=inpaint_resample_times, init_images=unet_init_images, skip_steps=unet_skip_steps,
                    cond_scale=unet_cond_scale, lowres_cond_img=lowres_cond_img, noise_scheduler=noise_scheduler,
                    pred_objective=pred_objective, dynamic_threshold=dynamic_threshold, use_tqdm=use_tqdm)
                outputs.append(img)
                if stop_at_unet_number is not None and stop_at_unet_number == unet_number:
                    break
            output_index = -1 if not return_all_outputs else slice(None)
            return outputs[output_index], lst_pred_noisy, lst_pred
        finally:
            self.train(was_training)

    def p_losses(self, unet, x_start, times, *, noise_scheduler, lowres_cond_img=None, cond_images=None, noise=None, pred_objective='noise',
                 p2_loss_weight_gamma=0., **kwargs):
        """imagen_pytorch3D.py:2277-2387 on the kernels (diffusioniqt_b200/train.py): returns (loss, pred, x_noisy, lowres_cond_img);
        `loss.backward()` runs the hand-written reverse pass and accumulates `.grad` on the U-Net's parameters."""
        from .train import p_losses
        if cond_images is not None:
            raise NotImplementedError("the training step does not implement cond_images")
        if getattr(self, 'use_lpips', False):
            raise NotImplementedError("the training step does not implement the LPIPS term (:2366-2381): it needs a pretrained network")
        return p_losses(self, unet, x_start, times, noise_scheduler=noise_scheduler, lowres_cond_img=lowres_cond_img, noise=noise,
                        pred_objective=pred_objective, p2_loss_weight_gamma=p2_loss_weight_gamma)

    def forward(self, images, lowres_img=None, unet=None, text_embeds=None, text_masks=None, unet_number=None, cond_images=None, **kwargs):
        """imagen_pytorch3D.py:2389-2442: one training evaluation of U-Net `unet_number` on a batch of high-field patches."""
        assert images.shape[-1] == images.shape[-2], f'the images you pass in must be a square, but received dimensions of {images.shape[2]}, {images.shape[-1]}'
        assert not (len(self.unets) > 1 and unet_number is None), \
            f'you must specify which unet you want trained, from a range of 1 to {len(self.unets)}, if you are training cascading DDPM (multiple unets)'
        unet_number = unet_number if unet_number is not None else 1
        assert self.only_train_unet_number is None or self.only_train_unet_number == unet_number, \
            f'you can only train on unet #{self.only_train_unet_number}'
        assert images.dtype == torch.float, f'images tensor needs to be floats but {images.dtype} dtype found instead'
        unet_index = unet_number - 1
        unet = unet if unet is not None else self.get_unet(unet_number)
        assert not isinstance(unet, NullUnet), 'null unet cannot and should not be trained'
        noise_scheduler = self.noise_schedulers[unet_index]
        p2_loss_weight_gamma = self.p2_loss_weight_gamma[unet_index]
        pred_objective = self.pred_objectives[unet_index]
        target_image_size = self.image_sizes[unet_index]
        b, c = images.shape[:2]
        assert c == self.channels and images.dim() == 5
        assert images.shape[-2] >= target_image_size and images.shape[-3] >= target_image_size
        if self.configs['Train']['batch_sample']:
            times = noise_scheduler.sample_random_times(1, device=images.device).repeat(b)
        else:
            times = noise_scheduler.sample_random_times(b, device=images.device)
        assert lowres_img is not None, 'lowres image must be provided'
        self.lowres_cond_img = lowres_img
        return self.p_losses(unet, images, times, cond_images=cond_images, noise_scheduler=noise_scheduler, lowres_cond_img=lowres_img,
                             pred_objective=pred_objective, p2_loss_weight_gamma=p2_loss_weight_gamma, **kwargs)
