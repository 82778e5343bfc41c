"""Host-side mirror of `ElucidatedImagen` (/root/reference/elucidated_imagen.py:75-700), sampling half.

The reference class cannot be constructed with, or call, the 3-D `Unet` of imagen_pytorch3D.py (SURVEY.md
Appendix C), so this is the working equivalent with the minimal adapter listed there: sigma is padded to 5-D, the U-Net
is called as `unet(c_in * x, <unused>, c_noise(sigma), lowres_cond_img=lr)`, and the low-res conditioning patch is not
noised (the U-Net was trained on the clean low-field patch, imagen_pytorch3D.py:2303-2304).  Constructor keywords,
`sample` / `one_unet_sample` signatures, the Karras schedule, churn, preconditioning and Heun correction follow
elucidated_imagen.py:75-106, :298-379, :382-532, :534-560.

One Heun step = two CUDA-graph replays:
    pass 0: edm_prepare (x_hat = x + churn noise, U-Net input = c_in x_hat) -> U-Net engine -> final conv fused with
            D(x_hat), slope d and the Euler step, which also writes the next U-Net input c_in(sigma_next) x_euler
    pass 1: U-Net engine -> final conv fused with D(x_euler), d' and the Heun update of the state
A device-side forward counter selects the rows of the per-forward constant table and of the FiLM table, so the two
graphs serve every step.  All sampler state stays fp32 on the device.
"""
from __future__ import annotations

from collections import namedtuple
from math import sqrt

import torch
import torch.nn.functional as F
from torch import nn

from . import lib as L
from .imagen import GaussianDiffusionContinuousTimes
from .unet import NullUnet, Unet, _cast_tuple

Hparams_fields = ['num_sample_steps', 'sigma_min', 'sigma_max', 'sigma_data', 'rho', 'P_mean', 'P_std', 'S_churn', 'S_tmin', 'S_tmax',
                  'S_noise']
Hparams = namedtuple('Hparams', Hparams_fields)

_EDM_ROW = 16


def _log(t, eps=1e-20):
    return torch.log(t.clamp(min=eps))


class _EdmState:
    def __init__(self):
        self.graphs = None


class ElucidatedImagen(nn.Module):
    def __init__(
        self,
        unets,
        *,
        image_sizes,
        text_encoder_name=None,
        text_embed_dim=None,
        channels=3,
        cond_drop_prob=0.1,
        random_crop_sizes=None,
        temporal_downsample_factor=1,
        lowres_sample_noise_level=0.2,
        per_sample_random_aug_noise_level=False,
        condition_on_text=False,       # the reference default (True) needs the T5 text path its 3-D U-Net does not have
        auto_normalize_img=True,
        dynamic_thresholding=True,
        dynamic_thresholding_percentile=0.95,
        only_train_unet_number=None,
        lowres_noise_schedule='linear',
        num_sample_steps=32,
        sigma_min=0.002,
        sigma_max=80,
        sigma_data=0.5,
        rho=7,
        P_mean=-1.2,
        P_std=1.2,
        S_churn=80,
        S_tmin=0.05,
        S_tmax=50,
        S_noise=1.003,
        clamp_range=(-1., 1.),         # not in the reference: its literal clamp(-1, 1) (:300, :527) assumes [-1, 1] data; z-score users pass (min_bound, inf)
    ):
        super().__init__()
        if condition_on_text:
            raise NotImplementedError("text conditioning: the 3-D U-Net of imagen_pytorch3D.py has no text path (SURVEY.md Appendix C)")
        self.only_train_unet_number = only_train_unet_number
        self.condition_on_text = False
        self.unconditional = True
        self.channels = channels
        unets = _cast_tuple(unets)
        num_unets = len(unets)
        self.random_crop_sizes = _cast_tuple(random_crop_sizes, num_unets)
        assert self.random_crop_sizes[0] is None, 'you should not need to randomly crop image during training for base unet, only for upsamplers'
        self.lowres_noise_schedule = GaussianDiffusionContinuousTimes(noise_schedule=lowres_noise_schedule)
        self.text_encoder_name = text_encoder_name
        self.text_embed_dim = text_embed_dim
        self.unets = nn.ModuleList([])
        self.unet_being_trained_index = -1
        for ind, one_unet in enumerate(unets):
            assert isinstance(one_unet, (Unet, NullUnet))
            # elucidated_imagen.py:152-158 also passes cond_on_text / text_embed_dim, which Unet.cast_model_parameters rejects
            one_unet = one_unet.cast_model_parameters(lowres_cond=not ind == 0, channels=self.channels, channels_out=self.channels)
            self.unets.append(one_unet)
        self.is_video = True            # volumes: sigma is padded 'b -> b 1 1 1 1'
        self.image_sizes = _cast_tuple(image_sizes)
        assert num_unets == len(self.image_sizes), f'you did not supply the correct number of u-nets ({len(self.unets)}) for resolutions {self.image_sizes}'
        self.sample_channels = _cast_tuple(self.channels, num_unets)
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
        self.input_image_range = (0. if auto_normalize_img else -1., 1.)
        self.dynamic_thresholding = _cast_tuple(dynamic_thresholding, num_unets)
        self.dynamic_thresholding_percentile = dynamic_thresholding_percentile
        temporal_downsample_factor = _cast_tuple(temporal_downsample_factor, num_unets)
        self.temporal_downsample_factor = temporal_downsample_factor
        assert temporal_downsample_factor[-1] == 1, 'downsample factor of last stage must be 1'
        hparams = [num_sample_steps, sigma_min, sigma_max, sigma_data, rho, P_mean, P_std, S_churn, S_tmin, S_tmax, S_noise]
        hparams = [_cast_tuple(hp, num_unets) for hp in hparams]
        self.hparams = [Hparams(*unet_hp) for unet_hp in zip(*hparams)]
        self.clamp_range = (float(clamp_range[0]), float(clamp_range[1]))
        self.register_buffer('_temp', torch.tensor([0.]), persistent=False)

        # ---- execution options of this implementation (not in the reference)
        self.use_cuda_graph = True
        self.noise_override = None     # tests: recorded tensors consumed instead of torch.randn, in the reference's draw order
        self.last_x_start = None       # x_start estimate of the last step (the reference keeps it for self-conditioning only)
        self.last_graph_launches = 0
        self._samplers = {}
        self.to(next(self.unets.parameters()).device)

    # ------------------------------------------------------------------ API parity helpers
    @property
    def device(self):
        return self._temp.device

    def force_unconditional_(self):
        self.condition_on_text = False
        self.unconditional = True

    def get_unet(self, unet_number):
        assert 0 < unet_number <= len(self.unets)
        return self.unets[unet_number - 1]

    def reset_unets_all_one_device(self, device=None):
        self.unets.to(device if device is not None else self.device)
        self.unet_being_trained_index = -1

    def load_state_dict(self, *args, **kwargs):
        self._samplers = {}
        for u in self.unets:
            if isinstance(u, Unet):
                u.invalidate_engines()
        return super().load_state_dict(*args, **kwargs)

    # ------------------------------------------------------------------ Table 1 of the paper (:308-324) and the schedule (:365-379)
    def c_skip(self, sigma_data, sigma):
        return (sigma_data ** 2) / (sigma ** 2 + sigma_data ** 2)

    def c_out(self, sigma_data, sigma):
        return sigma * sigma_data * (sigma_data ** 2 + sigma ** 2) ** -0.5

    def c_in(self, sigma_data, sigma):
        return 1 * (sigma ** 2 + sigma_data ** 2) ** -0.5

    def c_noise(self, sigma):
        return _log(sigma) * 0.25

    def sample_schedule(self, num_sample_steps, rho, sigma_min, sigma_max):
        N = num_sample_steps
        inv_rho = 1 / rho
        steps = torch.arange(num_sample_steps, device=self.device, dtype=torch.float32)
        sigmas = (sigma_max ** inv_rho + steps / (N - 1) * (sigma_min ** inv_rho - sigma_max ** inv_rho)) ** rho
        return F.pad(sigmas, (0, 1), value=0.)

    def threshold_x_start(self, x_start, dynamic_threshold=True):
        # :298-311
        if not dynamic_threshold:
            return x_start.clamp(*self.clamp_range)
        s = torch.quantile(x_start.reshape(x_start.shape[0], -1).abs(), self.dynamic_thresholding_percentile, dim=-1)
        s.clamp_(min=1.)
        s = s.reshape(-1, *((1,) * (x_start.dim() - 1)))
        return x_start.clamp(-s, s) / s

    # ------------------------------------------------------------------ per-forward constant table
    def _build_table(self, hp, sigma_min, sigma_max, skip_steps, clamp, dynamic_threshold):
        """One row per U-Net forward (2 per step; row 2i is the Euler pass at sigma_hat_i, row 2i+1 the Heun pass at
        sigma_{i+1}).  Scalars follow the reference's arithmetic: sigma / sigma_next / gamma become Python floats (:471),
        sigma_hat and the step sizes are float64 expressions, the preconditioning coefficients are fp32 tensor ops on
        `torch.full((b,), sigma)` (:340-346)."""
        dev = self.device
        sigmas = self.sample_schedule(hp.num_sample_steps, hp.rho, sigma_min, sigma_max)
        gammas = torch.where((sigmas >= hp.S_tmin) & (sigmas <= hp.S_tmax), min(hp.S_churn / hp.num_sample_steps, sqrt(2) - 1), 0.)
        sched = list(zip(sigmas[:-1].tolist(), sigmas[1:].tolist(), gammas[:-1].tolist()))[(skip_steps or 0):]
        lo, hi = self.clamp_range if (clamp and not dynamic_threshold) else (float('-inf'), float('inf'))
        rows, fwd_sigma = [], []
        for sigma, sigma_next, gamma in sched:
            sigma_hat = sigma + gamma * sigma
            for p, sg in ((0, sigma_hat), (1, sigma_next)):
                fwd_sigma.append(sg)
                rows.append([hp.S_noise, sqrt(sigma_hat ** 2 - sigma ** 2), 0., 0., 0., sg, sigma_next - sigma_hat, 0.5 * (sigma_next - sigma_hat),
                             0., lo, hi] + [0.] * (_EDM_ROW - 11))
        table = torch.tensor(rows, dtype=torch.float64).to(torch.float32).to(dev)
        sig = torch.tensor(fwd_sigma, dtype=torch.float64).to(torch.float32).to(dev)          # == torch.full((b,), sigma) per forward
        table[:, 2] = self.c_in(hp.sigma_data, sig)
        table[:, 3] = self.c_skip(hp.sigma_data, sig)
        table[:, 4] = self.c_out(hp.sigma_data, sig)
        table[0::2, 8] = table[1::2, 2]                                                       # Euler rows carry c_in(sigma_next)
        return table.contiguous(), self.c_noise(sig).contiguous(), sched, sigmas[0]

    # ------------------------------------------------------------------ the hot loop (:382-532)
    @torch.no_grad()
    def one_unet_sample(self, unet, shape, *, unet_number, clamp=True, dynamic_threshold=True, cond_scale=1., use_tqdm=True,
                        inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5, init_images=None, skip_steps=None,
                        sigma_min=None, sigma_max=None, lowres_cond_img=None, cond_images=None, **kwargs):
        if inpaint_images is not None or inpaint_masks is not None:
            raise NotImplementedError("inpainting (repaint) is not built for the Elucidated sampler")
        if cond_scale != 1:
            raise NotImplementedError("classifier-free guidance (cond_scale != 1) is not on the shipped sampling path")
        if unet.self_cond:
            raise NotImplementedError("self conditioning cannot run in the reference U-Net either (imagen_pytorch3D.py:1273-1286)")
        device = self.device
        if device.type != 'cuda':
            raise RuntimeError("ElucidatedImagen.sample runs only on a CUDA device (sm_100a kernels; there is no CPU fallback)")
        lib = L.load()
        hp = self.hparams[unet_number - 1]
        sigma_min = sigma_min if sigma_min is not None else hp.sigma_min
        sigma_max = sigma_max if sigma_max is not None else hp.sigma_max
        batch = shape[0]
        eng = unet.engine_for(batch, shape[2:], device)
        if eng.sub_f > 1:
            raise NotImplementedError("boundary mode with the Elucidated sampler")

        key = (tuple(hp), float(sigma_min), float(sigma_max), skip_steps or 0, bool(clamp), bool(dynamic_threshold), self.clamp_range)
        st = eng.sampler_cache.get(key)
        if st is None:
            st = _EdmState()
            st.table, st.c_noise, st.sched, st.init_sigma = self._build_table(hp, sigma_min, sigma_max, skip_steps, clamp, dynamic_threshold)
            st.fwd = torch.zeros(1, dtype=torch.int32, device=device)
            for name in ("x", "x_hat", "slope", "x0", "eps"):
                setattr(st, name, torch.empty(shape, dtype=torch.float32, device=device))
            eng.sampler_cache[key] = st
        inj = iter(self.noise_override) if self.noise_override is not None else None

        def draw(dst):
            if inj is not None:
                return dst.copy_(next(inj).to(device=device, dtype=torch.float32))
            return dst.normal_()

        draw(st.x)
        st.x.mul_(st.init_sigma)                                                    # :430-432
        if init_images is not None:
            st.x += init_images                                                     # :436-437
        eng.load_inputs(None, lowres_cond_img, cond_images)
        eng.set_condition(st.c_noise)                                               # time MLPs of every forward at once
        if getattr(st, "film_gen", None) != eng.film_gen:                           # the FiLM table moved: captured graphs read the old one
            st.graphs, st.film_gen = None, eng.film_gen
        eng.film_row_ptr, eng.film_stride_n = st.fwd.data_ptr(), 0
        st.fwd.zero_()
        count = st.x.numel()
        u = unet
        fused = not (clamp and dynamic_threshold)
        tab, fwd = st.table.data_ptr(), st.fwd.data_ptr()

        def final(pas):
            x = eng.last_act
            if fused:
                L.check(lib.diqt_final_conv_edm(x.ptr, x.ld, eng.ddtype, eng.conv_n, eng.level_vox[0] * (eng.n // eng.conv_n), x.c, u.channels_out,
                                                eng.w_final.data_ptr(), eng.b_final.data_ptr(), pas, tab, fwd, st.x_hat.data_ptr(),
                                                st.slope.data_ptr(), st.x.data_ptr(), st.x0.data_ptr(), eng.x_in.data_ptr(), 0, 0,
                                                L.current_stream()), "final_conv_edm")
            else:
                # dynamic thresholding takes a per-sample quantile (a sort): torch call between our kernels (like the DDPM path)
                eng.run_final(fused=False)
                row = st.table[int(st.fwd.item())]
                src = st.x_hat if pas == 0 else st.x
                st.x0.copy_(self.threshold_x_start(row[3] * src + row[4] * eng.pred, True))
                L.check(lib.diqt_edm_update(st.x0.data_ptr(), pas, tab, fwd, st.x_hat.data_ptr(), st.slope.data_ptr(), st.x.data_ptr(),
                                            eng.x_in.data_ptr(), count, L.current_stream()), "edm_update")
            L.check(lib.diqt_advance_step(fwd, L.current_stream()), "advance_step")

        def pass0():
            L.check(lib.diqt_edm_prepare(st.x.data_ptr(), st.eps.data_ptr(), tab, fwd, st.x_hat.data_ptr(), eng.x_in.data_ptr(), count,
                                         L.current_stream()), "edm_prepare")
            eng.run_body()
            final(0)

        def pass1():
            eng.run_body()
            final(1)

        use_graph = self.use_cuda_graph and fused
        if use_graph and st.graphs is None:
            saved = st.x.clone()
            st.eps.zero_()
            s = torch.cuda.Stream(device=device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):                                              # warm-up outside capture (lazy loading, func attributes)
                pass0()
                pass1()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize(device)
            st.x.copy_(saved)
            st.fwd.zero_()
            graphs, launches = [], 0
            for fn in (pass0, pass1):
                g = torch.cuda.CUDAGraph()
                before = L.launch_count()
                with torch.cuda.graph(g):
                    fn()
                launches += L.launch_count() - before
                graphs.append(g)
            st.graphs, st.graph_launches = graphs, launches
        for ind, (sigma, sigma_next, gamma) in enumerate(st.sched):
            draw(st.eps)                                                            # :476, drawn every step
            if use_graph:
                st.graphs[0].replay()
            else:
                pass0()
            if sigma_next != 0:                                                     # second-order correction (:500-516)
                if use_graph:
                    st.graphs[1].replay()
                else:
                    pass1()
            else:
                L.check(lib.diqt_advance_step(fwd, L.current_stream()), "advance_step")
        self.last_graph_launches = getattr(st, "graph_launches", 0)
        self.last_x_start = st.x0
        images = st.x.clone()
        L.check(lib.diqt_clamp(images.data_ptr(), images.numel(), self.clamp_range[0], self.clamp_range[1], L.current_stream()), "clamp")   # :527
        return self.unnormalize_img(images)

    @torch.no_grad()
    def sample(self, texts=None, text_masks=None, text_embeds=None, cond_images=None, inpaint_images=None, inpaint_masks=None,
               inpaint_resample_times=5, init_images=None, skip_steps=None, sigma_min=None, sigma_max=None, video_frames=None,
               batch_size=1, cond_scale=1., lowres_sample_noise_level=None, start_at_unet_number=1, start_image_or_video=None,
               stop_at_unet_number=None, return_all_unet_outputs=False, return_pil_images=False, use_tqdm=True, device=None):
        """Signature of elucidated_imagen.py:534-560; returns the last U-Net's output (or all of them)."""
        was_training = self.training
        self.eval()
        try:
            device = device if device is not None else self.device
            self.reset_unets_all_one_device(device=device)
            assert texts is None and text_embeds is None and text_masks is None, 'imagen specified not to be conditioned on text, yet it is presented'
            assert not ((inpaint_images is None) ^ (inpaint_masks is None)), 'inpaint images and masks must be both passed in to do inpainting'
            assert not return_pil_images, 'volumes cannot be converted to PIL images'
            if cond_images is not None and cond_images.dtype == torch.uint8:
                cond_images = cond_images / 255
            num_unets = len(self.unets)
            cond_scale = _cast_tuple(cond_scale, num_unets)
            init_images = [self.normalize_img(i) if i is not None else None for i in _cast_tuple(init_images, num_unets)]
            skip_steps = _cast_tuple(skip_steps, num_unets)
            sigma_min = _cast_tuple(sigma_min, num_unets)
            sigma_max = _cast_tuple(sigma_max, num_unets)
            img = None
            if start_at_unet_number > 1:
                assert start_at_unet_number <= num_unets, 'must start a unet that is less than the total number of unets'
                assert stop_at_unet_number is None or start_at_unet_number <= stop_at_unet_number
                assert start_image_or_video is not None, 'starting image or video must be supplied if only doing upscaling'
                img = start_image_or_video            # patches already have the target size (no resize, as in Imagen.sample :2244)
            outputs = []
            for unet_number, unet, image_size, dynamic_threshold, unet_cond_scale, unet_init_images, unet_skip_steps, unet_sigma_min, unet_sigma_max in zip(
                    range(1, num_unets + 1), self.unets, self.image_sizes, self.dynamic_thresholding, cond_scale, init_images, skip_steps,
                    sigma_min, sigma_max):
                if unet_number < start_at_unet_number:
                    continue
                assert not isinstance(unet, NullUnet), 'cannot sample from null unet'
                lowres_cond_img = None
                if unet.lowres_cond:
                    assert img is not None, 'low resolution conditioning image must be present'
                    # the reference noises this image at level 0.2 (:651-657) for a U-Net input the 3-D model lacks: left clean
                    lowres_cond_img = self.normalize_img(img.to(device=device, dtype=torch.float32))
                shape = (batch_size, self.channels, image_size, image_size, image_size)
                img = self.one_unet_sample(unet, shape, unet_number=unet_number, cond_images=cond_images, inpaint_images=inpaint_images,
                                           inpaint_masks=inpaint_masks, inpaint_resample_times=inpaint_resample_times,
                                           init_images=unet_init_images, skip_steps=unet_skip_steps, sigma_min=unet_sigma_min,
                                           sigma_max=unet_sigma_max, cond_scale=unet_cond_scale, lowres_cond_img=lowres_cond_img,
                                           dynamic_threshold=dynamic_threshold, use_tqdm=use_tqdm)
                outputs.append(img)
                if stop_at_unet_number is not None and stop_at_unet_number == unet_number:
                    break
            return outputs[-1] if not return_all_unet_outputs else outputs
        finally:
            self.train(was_training)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("training (ElucidatedImagen.forward, elucidated_imagen.py:704-846) is outside the sampling hot path")
