"""DDIM sampler with inversion guidance -- the VFace denoising loop -- on the vface_b200 kernels.

Host-side mirror of REFace/ldm/models/diffusion/ddim_w_inv.py::DDIMSampler (:141-738): the public
methods, their argument names and return shapes are the reference's, so
scripts/VFace_inference_batch.py (:531 ddim_invert, :580 sample) runs unchanged:

    make_schedule :155-184   sample :186-252   ddim_sampling :254-355   ddim_invert :360-490
    p_sample_ddim :564-617   p_sample_ddim_with_inverse :621-738

Semantics kept: the live step is p_sample_ddim_with_inverse; the UNet
batch is [uncond(x,uc) ; cond(x,c) ; recon(ddim_inv_t, target_cond)]; hooks are switched off on all
16 attn1 modules and on (fusion="flow_fix", split 0.8, alpha 0.8) for input_blocks only; noise is
drawn twice per step; (samples, intermediates) is returned.

Re-designed: the hook configuration is loop-invariant and registered once; CFG + DDIM update is one
kernel fed by host-side fp32 scalars (no torch.full / .item() syncs); inversion latents are read
once per sample() call (or handed over in memory from ddim_invert) instead of one torch.load per
step; with `elide_dead_recon=True` (extension, off by default) the output-dead recon branch
 is not computed.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ...modules.diffusionmodules.util import make_ddim_sampling_parameters, make_ddim_timesteps, noise_like
from ..pnp_utils import register_spa_attn_injection
from .... import frame_shard, ops


def load_ddim_latents_at_t(t, ddim_latents_path):
    path = os.path.join(ddim_latents_path, f"ddim_latents_{t}.pt")
    assert os.path.exists(path), f"Missing latents at t {t} path {path}"
    return torch.load(path)


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", elide_dead_recon=False, cuda_graphs=None, **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.elide_dead_recon = bool(elide_dead_recon)
        self.last_inversion = None          # {timestep: latents} of the most recent ddim_invert
        self._inv_cache = None
        # Extension: one CUDA graph per denoising step, captured the first time a step of a given
        # (schedule position, shapes) is run and replayed for every later batch of the clip -- the ~370 launches of a step
        # become one.  Off by default (VF_CUDA_GRAPH=1 or cuda_graphs=True); single-rank only.
        self.cuda_graphs = bool(int(os.environ.get("VF_CUDA_GRAPH", "0"))) if cuda_graphs is None else bool(cuda_graphs)
        self._graphs = {}
        self._graph_pool = None
        self._graph_flow = None

    # -- buffers / schedule ----------------------------------------------------------------------
    def register_buffer(self, name, attr):
        if isinstance(attr, torch.Tensor):
            dev = self.model.betas.device
            if attr.device != dev:
                attr = attr.to(dev)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.model.betas.device)

        self.register_buffer('betas', to_torch(self.model.betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(self.model.alphas_cumprod_prev))
        acp_cpu = alphas_cumprod.detach().float().cpu()
        self.register_buffer('sqrt_alphas_cumprod', to_torch(torch.sqrt(acp_cpu)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(torch.sqrt(1. - acp_cpu)))

        ddim_sigmas, ddim_alphas, ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=acp_cpu, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        self.register_buffer('ddim_sigmas', ddim_sigmas)
        self.register_buffer('ddim_alphas', ddim_alphas)
        self.register_buffer('ddim_alphas_prev', ddim_alphas_prev)
        self.register_buffer('ddim_sqrt_one_minus_alphas', torch.sqrt(1. - ddim_alphas.cpu()))
        # Host-side fp32 scalars for the fused update kernel: what torch.full((b,1,1,1), table[index])
        # would have produced (ddim_w_inv.py:679-682), without a device round trip per step.
        f32 = lambda a: np.asarray(torch.as_tensor(a).detach().cpu().numpy(), dtype=np.float64).astype(np.float32)
        self._host_tables = dict(a_t=f32(ddim_alphas), a_prev=f32(ddim_alphas_prev), sigma=f32(ddim_sigmas),
                                 s1m=f32(self.ddim_sqrt_one_minus_alphas), acp=f32(acp_cpu))

    # -- public sampling entry -------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, target_conditioning=None, inverse_results_dir=None,
               callback=None, normals_sequence=None, img_callback=None, quantize_x0=False, eta=0., mask=None,
               x0=None, temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
               verbose=True, flow=None, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, src_im=None, tar=None, **kwargs):
        if conditioning is not None and not isinstance(conditioning, dict) and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        return self.ddim_sampling(conditioning, size, target_conditioning=target_conditioning,
                                  inverse_results_dir=inverse_results_dir, callback=callback,
                                  img_callback=img_callback, quantize_denoised=quantize_x0, mask=mask, x0=x0,
                                  ddim_use_original_steps=False, noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  flow=flow, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, src_im=src_im, **kwargs)

    def _register_hooks(self, flow):
        """The effective configuration of ddim_w_inv.py:300-308: off everywhere, then flow_fix on the
        six input-block attn1 modules.  Loop-invariant, so registered once per sampling call."""
        register_spa_attn_injection(self, 1, switch_on=False, input_blocks=True, middle_block=True, output_blocks=True,
                                    attn_component="attn1", flow=flow, chunks=3,
                                    block_indices=[0, 1, 2, 3, 4, 5, 6, 7, 8], fusion="flow_fix",
                                    split_ratio_fft=0.8, alpha=0.8, _elide_recon=self.elide_dead_recon)
        register_spa_attn_injection(self, 1, switch_on=True, input_blocks=True, middle_block=False, output_blocks=False,
                                    attn_component="attn1", flow=flow, chunks=3,
                                    block_indices=[0, 1, 2, 3, 4, 5, 6, 7, 8], fusion="flow_fix",
                                    split_ratio_fft=0.8, alpha=0.8, _elide_recon=self.elide_dead_recon)

    def _prefetch_inversion(self, inverse_results_dir, steps, device):
        """All per-timestep inversion latents on the device, once (reference: one torch.load per step,
        ddim_w_inv.py:22-26, :628).  `inverse_results_dir` may be the reference's directory of
        ddim_latents_{t}.pt files or a {t: tensor} dict handed over from ddim_invert."""
        cache = {}
        shard = frame_shard.current()
        for t in steps:
            t = int(t)
            if isinstance(inverse_results_dir, dict):
                lat = inverse_results_dir[t]
            else:
                lat = load_ddim_latents_at_t(t, inverse_results_dir)
            lat = lat.to(device=device, dtype=torch.float32, non_blocking=True)
            if shard is not None and lat.shape[0] == shard.total_frames and shard.world_size > 1:
                lat = shard.take(lat)
            cache[t] = lat
        return cache

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, target_conditioning=None, inverse_results_dir=None, x_T=None,
                      ddim_use_original_steps=False, callback=None, timesteps=None, quantize_denoised=False,
                      mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1., noise_dropout=0.,
                      score_corrector=None, flow=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, src_im=None, **kwargs):
        if ddim_use_original_steps or timesteps is not None:
            raise NotImplementedError("ddim_use_original_steps / timesteps subsets are not used by the VFace scripts")
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        timesteps = self.ddim_timesteps
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]

        graphs = (self.cuda_graphs and target_conditioning is not None and torch.device(device).type == "cuda"
                  and not (frame_shard.current() is not None and frame_shard.current().world_size > 1)
                  and mask is None and not callback and not img_callback
                  and unconditional_conditioning is not None and unconditional_guidance_scale != 1.)
        if graphs and flow is not None:
            # the hooks keep the flow object they were registered with: give them a buffer whose address outlives the call
            fl = ops._as_flow(flow, device)
            if self._graph_flow is None or self._graph_flow.shape != fl.shape:
                self._graph_flow = torch.empty_like(fl)
                self._graphs.clear()
            self._graph_flow.copy_(fl)
            flow = self._graph_flow
        self._register_hooks(flow)
        self._inv_cache = None
        if target_conditioning is not None:
            self._inv_cache = self._prefetch_inversion(inverse_results_dir, time_range, device)

        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            common = dict(index=index, use_original_steps=False, quantize_denoised=quantize_denoised,
                          temperature=temperature, noise_dropout=noise_dropout, score_corrector=score_corrector,
                          corrector_kwargs=corrector_kwargs,
                          unconditional_guidance_scale=unconditional_guidance_scale,
                          unconditional_conditioning=unconditional_conditioning, **kwargs)
            if graphs:
                outs = self._graphed_step(img, cond, int(step), index, target_conditioning, flow, common)
            elif target_conditioning is not None:
                outs = self.p_sample_ddim_with_inverse(img, cond, ts, target_conditioning=target_conditioning,
                                                       inverse_results_dir=inverse_results_dir, src_start=None,
                                                       flow=flow, _step=int(step), **common)
            else:
                outs = self.p_sample_ddim(img, cond, ts, **common)
            img, pred_x0 = outs
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        self._inv_cache = None
        return img, intermediates

    # -- one reverse step as a CUDA graph (extension) ------------------------------------------------
    def _graphed_step(self, img, cond, step, index, target_conditioning, flow, common):
        """p_sample_ddim_with_inverse for schedule position `index` through a CUDA graph: captured on first use (after one
        eager warm-up run on a side stream, which also creates the library plans and function attributes), replayed
        afterwards with the inputs copied into the graph's static buffers.  The host-side scalars of the update
        (a_t, a_prev, sigma_t, sqrt(1-a_t), the guidance scale) are kernel arguments, so they are part of the key."""
        tmk = common.get("test_model_kwargs")
        if tmk is None:
            raise NotImplementedError("cuda_graphs needs test_model_kwargs (the VFace scripts' calling convention)")
        uc = common["unconditional_conditioning"]
        inv = self._inv_cache[step]
        live = dict(img=img, cond=cond, tc=target_conditioning, uc=uc, ii=tmk["inpaint_image"], im=tmk["inpaint_mask"], inv=inv)
        tb = self._host_tables
        key = (index, step, float(tb["a_t"][index]), float(tb["a_prev"][index]), float(tb["sigma"][index]),
               float(common["unconditional_guidance_scale"]), float(common.get("temperature", 1.)), self.elide_dead_recon,
               tuple((k, tuple(v.shape), v.dtype) for k, v in live.items()), None if flow is None else flow.data_ptr(),
               next(self.model.model.diffusion_model.parameters()).dtype)
        entry = self._graphs.get(key)
        if entry is None:
            static = {k: v.clone() for k, v in live.items()}

            def run():
                kw = dict(common)
                kw.pop("index", None)
                kw["index"] = index
                kw["test_model_kwargs"] = dict(inpaint_image=static["ii"], inpaint_mask=static["im"])
                kw["unconditional_conditioning"] = static["uc"]
                saved = self._inv_cache
                self._inv_cache = {step: static["inv"]}
                try:
                    ts = torch.full((static["img"].shape[0],), step, device=static["img"].device, dtype=torch.long)
                    return self.p_sample_ddim_with_inverse(static["img"], static["cond"], ts, target_conditioning=static["tc"],
                                                           inverse_results_dir=None, src_start=None, flow=flow, _step=step, **kw)
                finally:
                    self._inv_cache = saved

            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                run()                                   # warm-up: library plans, function attributes, allocator
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(g, pool=self._graph_pool):
                out = run()
            entry = self._graphs[key] = (g, static, out)
        g, static, out = entry
        for k, v in live.items():
            static[k].copy_(v)
        g.replay()
        return out[0].clone(), out[1].clone()

    # -- one reverse step --------------------------------------------------------------------------
    @staticmethod
    def _extra_channels(kwargs):
        if 'test_model_kwargs' in kwargs:
            k = kwargs['test_model_kwargs']
            return torch.cat([k['inpaint_image'], k['inpaint_mask']], dim=1)
        if 'rest' in kwargs:
            return kwargs['rest']
        raise Exception("kwargs must contain either 'test_model_kwargs' or 'rest' key")

    @staticmethod
    def _reject_unsupported(use_original_steps, quantize_denoised, noise_dropout, score_corrector):
        if use_original_steps or quantize_denoised or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError("use_original_steps / quantize_denoised / score_corrector / noise_dropout "
                                      "are not used by the VFace scripts")

    def _update(self, x_lat, e_uncond, e_cond, index, scale, temperature, repeat_noise, draws):
        """CFG + DDIM update in one kernel (ddim_w_inv.py:666, :679-700).  `draws` noise tensors are
        drawn per step like the reference (2 in p_sample_ddim_with_inverse, :697 and :704) so the RNG
        stream stays aligned; only the first is used."""
        tb = self._host_tables
        sigma = float(tb['sigma'][index])
        noise = None
        for j in range(draws):
            nz = noise_like(x_lat.shape, x_lat.device, repeat_noise)
            if j == 0 and sigma != 0.0:
                noise = nz if temperature == 1. else nz * temperature
        return ops.ddim_cfg_step(x_lat.contiguous(), e_uncond, e_cond, float(tb['a_t'][index]), float(tb['a_prev'][index]),
                                 sigma, float(tb['s1m'][index]), float(scale), noise)

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        self._reject_unsupported(use_original_steps, quantize_denoised, noise_dropout, score_corrector)
        x_lat = x if x.shape[1] == 4 else x[:, :4]
        x_full = torch.cat([x, self._extra_channels(kwargs)], dim=1)
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_t = self.model.apply_model(x_full, t, c)
            e_uncond, scale = e_t, 1.0
        else:
            x_in = torch.cat([x_full] * 2)
            t_in = torch.cat([t] * 2)
            c_in = torch.cat([unconditional_conditioning, c])
            e_uncond, e_t = self.model.apply_model(x_in, t_in, c_in).chunk(2)
            scale = unconditional_guidance_scale
        return self._update(x_lat, e_uncond, e_t, index, scale, temperature, repeat_noise, draws=1)

    @torch.no_grad()
    def p_sample_ddim_with_inverse(self, x, c, t, index, target_conditioning=None, inverse_results_dir=None,
                                   repeat_noise=False, src_start=None, use_original_steps=False,
                                   quantize_denoised=False, temperature=1., noise_dropout=0., score_corrector=None,
                                   corrector_kwargs=None, unconditional_guidance_scale=1., flow=None,
                                   unconditional_conditioning=None, _step=None, **kwargs):
        self._reject_unsupported(use_original_steps, quantize_denoised, noise_dropout, score_corrector)
        if src_start is not None:
            raise NotImplementedError("src_start is always None in the VFace scripts (ddim_w_inv.py:331)")
        extra = self._extra_channels(kwargs)
        x_full = torch.cat([x, extra], dim=1)
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_t = self.model.apply_model(x_full, t, c)
            return self._update(x, e_t, e_t, index, 1.0, temperature, repeat_noise, draws=2)
        if self.elide_dead_recon:
            x_in = torch.cat([x_full, x_full])
            t_in = torch.cat([t] * 2)
            c_in = torch.cat([unconditional_conditioning, c])
            e_uncond, e_t = self.model.apply_model(x_in, t_in, c_in).chunk(2)
        else:
            step = int(t[0].item()) if _step is None else _step
            if self._inv_cache is not None and step in self._inv_cache:
                ddim_inv_t = self._inv_cache[step]
            elif isinstance(inverse_results_dir, dict):
                ddim_inv_t = inverse_results_dir[step].to(x.device)
            else:
                ddim_inv_t = load_ddim_latents_at_t(step, inverse_results_dir).to(x.device)
            inv_full = torch.cat([ddim_inv_t.to(x.dtype), extra], dim=1)
            x_in = torch.cat([x_full, x_full, inv_full], dim=0)
            t_in = torch.cat([t] * 3)
            c_in = torch.cat([unconditional_conditioning, c, target_conditioning], dim=0)
            e_uncond, e_t, _e_recon = self.model.apply_model(x_in, t_in, c_in).chunk(3)
        return self._update(x, e_uncond, e_t, index, unconditional_guidance_scale, temperature, repeat_noise, draws=2)

    # -- inversion (pre-step) -------------------------------------------------------
    @torch.no_grad()
    def ddim_invert(self, x, cond, S, shape, eta=0., unconditional_guidance_scale=1.,
                    unconditional_conditioning=None, inverse_dir=None, batch_size=6, src_lm=None, tar_lm=None,
                    **kwargs):
        """Forward DDIM over the 2B batch [target latents ; source latents] with hooks off; the target
        half of every step is kept in `self.last_inversion[step]` and, if `inverse_dir` is given,
        written as ddim_latents_{step}.pt like the reference (:464-486)."""
        device = x.device
        b = x.shape[0]
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=False)
        timesteps = self.ddim_timesteps
        intermediates = {'x_inter': [x]}
        register_spa_attn_injection(self, 1, switch_on=False, input_blocks=True, middle_block=True, output_blocks=True,
                                    attn_component="attn1", chunks=3)
        extra = self._extra_channels(kwargs)
        acp = self._host_tables['acp']
        stride = 1000 // len(timesteps)
        saved = {}
        x = x.to(torch.float32)
        for i, step in enumerate(timesteps):
            step = int(step)
            ts = torch.full((b,), step, device=device, dtype=torch.long)
            x_full = torch.cat([x, extra], dim=1)
            if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
                e_t = self.model.apply_model(x_full, ts, cond)
                e_uncond, scale = None, 1.0
            else:
                x_in = torch.cat([x_full] * 2)
                t_in = torch.cat([ts] * 2)
                c_in = torch.cat([unconditional_conditioning, cond])
                e_uncond, e_t = self.model.apply_model(x_in, t_in, c_in).chunk(2)
                scale = unconditional_guidance_scale
            a_next = float(acp[step])
            a_cur = float(acp[max(0, step - stride)])
            x = ops.ddim_invert_step(x.contiguous(), e_t.contiguous(), a_cur, a_next,
                                     e_uncond=None if e_uncond is None else e_uncond.contiguous(), cfg_scale=scale)
            intermediates['x_inter'].append(x)
            saved[step] = x[:batch_size].detach().clone()
            if inverse_dir is not None:
                torch.save(saved[step], os.path.join(inverse_dir, f"ddim_latents_{step}.pt"))
        self.last_inversion = saved
        return x, intermediates
