"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden            (build container only: needs /root/reference)

The reference ships no golden vectors, so these outputs of the reference itself
are what pins the oracle port (tests/test_oracle_vs_golden.py) and, through it, the CUDA path.
Inputs are regenerated from seeds by the tests (vface_b200.synth + torch CPU generators, same torch
build in the container and on the GPU box: the versions are recorded in every file); only outputs and
small inputs are stored.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
META = dict(torch=torch.__version__, numpy=np.__version__, reference="Sanoojan/VFace REFace/ (unmodified, CPU fp32)")

SMALL_UNET = dict(model_channels=32, num_heads=2)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), meta=np.array(repr(META)), **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def golden_fsai():
    from scripts.face_swap_utils import combine_fft_high_low
    out = {}
    for d, ratios in ((320, (0.8, 0.5, 0.25)), (640, (0.8,)), (1280, (0.8,))):
        g = torch.Generator().manual_seed(100 + d)
        donor = torch.randn(2, 8, d, generator=g)
        dst = torch.randn(2, 8, d, generator=g)
        out[f"donor_{d}"] = donor.numpy()
        out[f"dst_{d}"] = dst.numpy()
        for r in ratios:
            with quiet():
                out[f"out_{d}_{int(r * 100)}"] = combine_fft_high_low(donor, dst, split_ratio=r).numpy()
    save("fsai.npz", **out)


def make_flows(kind, n, hw, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "smooth":
        return [torch.randn(1, 2, hw, hw, generator=g) * 3.0 for _ in range(n)]
    if kind == "integer":
        return [torch.randint(-5, 6, (1, 2, hw, hw), generator=g).float() for _ in range(n)]
    return [torch.randn(1, 2, hw, hw, generator=g) * 60.0 for _ in range(n)]


def golden_warp():
    from scripts.temporal_flow import align_by_flow, warp_image
    hw, c, frames = 64, 8, 3
    out = {}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(frames, c, hw, hw, generator=g)
    out["x"] = x.numpy()
    # one-hot probes: channel j of `cols` is 1 in column j, of `rows` in row j.  out[j, y, x] != 0
    # <=> column (row) j is a bilinear tap of pixel (y, x) -- recovers the reference's floor indices.
    cols = torch.zeros(1, hw, hw, hw)
    rows = torch.zeros(1, hw, hw, hw)
    for j in range(hw):
        cols[0, j, :, j] = 1.0
        rows[0, j, j, :] = 1.0
    for kind in ("smooth", "integer", "far"):
        flows = make_flows(kind, frames - 1, hw, 21)
        out[f"flow_{kind}"] = torch.cat(flows).numpy()
        out[f"out_{kind}"] = align_by_flow(x, flow=flows, alpha=0.8).numpy()
        wc = warp_image(cols, flows[0])[0]           # (hw, hw, hw): [column j, y, x]
        wr = warp_image(rows, flows[0])[0]
        out[f"x0_{kind}"] = (wc != 0).float().argmax(dim=0).to(torch.int16).numpy()   # first tap column
        out[f"y0_{kind}"] = (wr != 0).float().argmax(dim=0).to(torch.int16).numpy()
        out[f"ntaps_x_{kind}"] = (wc != 0).sum(dim=0).to(torch.int16).numpy()
        out[f"ntaps_y_{kind}"] = (wr != 0).sum(dim=0).to(torch.int16).numpy()
    save("warp.npz", **out)


def golden_attn_hooks():
    """The patched attn1 forward (pnp_utils.py:92-288) on a 2-head d=40 CrossAttention, N = 4096."""
    from ldm.modules.attention import CrossAttention
    from ldm.models.pnp_utils import register_spa_attn_injection
    torch.manual_seed(0)
    attn = CrossAttention(query_dim=80, heads=2, dim_head=40)
    g = torch.Generator().manual_seed(5)
    for p in attn.parameters():
        p.data = torch.randn(p.shape, generator=g) * (p.shape[-1] ** -0.5 if p.dim() > 1 else 0.02)

    class Holder:                                            # model.model.model.diffusion_model.input_blocks
        pass
    blk = torch.nn.Module()
    blk.attn1 = attn
    unet = torch.nn.Module()
    unet.input_blocks = torch.nn.ModuleList([blk])
    unet.middle_block = torch.nn.ModuleList([])
    unet.output_blocks = torch.nn.ModuleList([])
    h = Holder(); h.model = Holder(); h.model.model = Holder(); h.model.model.diffusion_model = unet
    B, N = 2, 4096
    x = torch.randn(3 * B, N, 80, generator=g)
    flows = make_flows("smooth", B - 1, 64, 33)
    out = {"flow": torch.cat(flows).numpy(), "params": np.array([0])}
    for name, p in attn.state_dict().items():
        out["w_" + name] = p.numpy()
    out["x_rows"] = x[:, ::32].numpy()
    for mode, kw in (("off", dict(switch_on=False, fusion="flow_fix")),
                     ("replace", dict(switch_on=True, fusion="replace")),
                     ("fft", dict(switch_on=True, fusion="fft", split_ratio_fft=0.8)),
                     ("flow_fix", dict(switch_on=True, fusion="flow_fix", split_ratio_fft=0.8, alpha=0.8, flow=flows))):
        with quiet():
            register_spa_attn_injection(h, 1, input_blocks=True, output_blocks=False, middle_block=False,
                                        attn_component="attn1", chunks=3, **kw)
            with torch.no_grad():
                y = attn.forward(x.clone())
        out[f"out_{mode}"] = y[:, ::32].numpy()              # every 32nd token row
    save("attn_hooks.npz", **out)


def golden_schedule():
    unet = torch.nn.Identity()
    out = {}
    for S in (10, 50):
        for eta in (0.0, 0.5):
            s = rh.build_reference_sampler(unet)
            with quiet():
                s.make_schedule(S, ddim_eta=eta, verbose=False)
            tag = f"S{S}_eta{int(eta * 10)}"
            out[f"timesteps_{tag}"] = np.asarray(s.ddim_timesteps)
            for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas"):
                out[f"{k}_{tag}"] = np.asarray(torch.as_tensor(getattr(s, k)).numpy(), dtype=np.float64).astype(np.float32)
    out["alphas_cumprod"] = rh.LatentDiffusionStub(unet).alphas_cumprod.numpy()
    save("schedule.npz", **out)


def golden_sampler_small():
    from vface_b200 import synth
    from . import kernels as ok
    ref = rh.build_reference_unet(SMALL_UNET)
    sd = synth.synth_state_dict(ref.state_dict(), seed=1)
    ref.load_state_dict(sd)
    B, S = 2, 10         # BASELINE.json config 1: DDIM 10 steps
    steps = ok.make_schedule(S)["ddim_timesteps"]
    out = {}
    for kind in ("smooth", "integer"):
        clip = synth.synth_clip(B, steps=steps, flow_kind=kind)
        sampler = rh.build_reference_sampler(ref)
        d = tempfile.mkdtemp()
        for t, v in clip["inversion"].items():
            torch.save(v, os.path.join(d, f"ddim_latents_{t}.pt"))
        xs = []
        with quiet():
            samples, inter = sampler.sample(
                S=S, batch_size=B, shape=(4, 64, 64), conditioning=clip["c"], target_conditioning=clip["target_cond"],
                inverse_results_dir=d, x_T=clip["x_T"], flow=clip["flow"], unconditional_guidance_scale=3.0,
                unconditional_conditioning=clip["uc"], eta=0.0, verbose=False, log_every_t=1,
                test_model_kwargs=dict(inpaint_image=clip["inpaint_image"], inpaint_mask=clip["inpaint_mask"]))
        out[f"samples_{kind}"] = samples.numpy()
        out[f"x_inter_{kind}"] = torch.stack(inter["x_inter"][1:]).numpy()
        out[f"pred_x0_{kind}"] = torch.stack(inter["pred_x0"][1:]).numpy()
    # ddim_invert on the same UNet (hooks off, 2B batch, no CFG): x_T and the saved target halves
    clip = synth.synth_clip(2 * B)
    sampler = rh.build_reference_sampler(ref)
    d = tempfile.mkdtemp()
    with quiet():
        xT, _ = sampler.ddim_invert(x=clip["x_T"], cond=clip["c"], S=S, shape=(4, 64, 64), eta=0.0, inverse_dir=d,
                                    batch_size=B, test_model_kwargs=dict(inpaint_image=clip["inpaint_image"],
                                                                         inpaint_mask=clip["inpaint_mask"]))
    out["invert_xT"] = xT.numpy()
    for t in steps:
        out[f"invert_saved_{int(t)}"] = torch.load(os.path.join(d, f"ddim_latents_{int(t)}.pt")).numpy()
    save("sampler_small.npz", **out)


def golden_sampler_full():
    """The FULL-SIZE UNet (project_ffhq.yaml, 859.5 M parameters) through the reference sampler: 2 frames, DDIM 5
    steps, CFG 3.0, hooks on (FSAI on the six input-block attn1 modules, flow warp on the two 64x64 ones) -- the
    per-step latents the `<= 1e-2 relative L2` bound of BASELINE.json is stated for.  ~3 minutes on 8 cores."""
    from vface_b200 import synth
    from . import kernels as ok
    ref = rh.build_reference_unet()
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), seed=1))
    B, S = 2, 5
    steps = ok.make_schedule(S)["ddim_timesteps"]
    clip = synth.synth_clip(B, steps=steps, flow_kind="smooth")
    sampler = rh.build_reference_sampler(ref)
    d = tempfile.mkdtemp()
    for t, v in clip["inversion"].items():
        torch.save(v, os.path.join(d, f"ddim_latents_{t}.pt"))
    with quiet():
        samples, inter = sampler.sample(
            S=S, batch_size=B, shape=(4, 64, 64), conditioning=clip["c"], target_conditioning=clip["target_cond"],
            inverse_results_dir=d, x_T=clip["x_T"], flow=clip["flow"], unconditional_guidance_scale=3.0,
            unconditional_conditioning=clip["uc"], eta=0.0, verbose=False, log_every_t=1,
            test_model_kwargs=dict(inpaint_image=clip["inpaint_image"], inpaint_mask=clip["inpaint_mask"]))
    save("sampler_full.npz", samples=samples.numpy(), x_inter=torch.stack(inter["x_inter"][1:]).numpy(),
         pred_x0=torch.stack(inter["pred_x0"][1:]).numpy())


def golden_sampler_full_s10():
    """BASELINE.json configs[0] at reduced frame count: the FULL-SIZE UNet through the reference sampler, DDIM 10 steps,
    2 frames, CFG 3.0, hooks on (integer-valued flow: the floor indices of the warp sit exactly on pixel centres).
    ~5 minutes on 8 cores; only the per-step latents are kept (fp16-rounded differences would defeat the purpose, so
    they stay fp32: 1.3 MB)."""
    from vface_b200 import synth
    from . import kernels as ok
    ref = rh.build_reference_unet()
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), seed=1))
    B, S = 2, 10
    steps = ok.make_schedule(S)["ddim_timesteps"]
    clip = synth.synth_clip(B, steps=steps, flow_kind="integer")
    sampler = rh.build_reference_sampler(ref)
    d = tempfile.mkdtemp()
    for t, v in clip["inversion"].items():
        torch.save(v, os.path.join(d, f"ddim_latents_{t}.pt"))
    with quiet():
        samples, inter = sampler.sample(
            S=S, batch_size=B, shape=(4, 64, 64), conditioning=clip["c"], target_conditioning=clip["target_cond"],
            inverse_results_dir=d, x_T=clip["x_T"], flow=clip["flow"], unconditional_guidance_scale=3.0,
            unconditional_conditioning=clip["uc"], eta=0.0, verbose=False, log_every_t=1,
            test_model_kwargs=dict(inpaint_image=clip["inpaint_image"], inpaint_mask=clip["inpaint_mask"]))
    save("sampler_full_s10.npz", x_inter=torch.stack(inter["x_inter"][1:]).numpy())


def golden_sampler_full_8f():
    """BASELINE.json configs[0] at its full frame count: 8 frames, DDIM 10 steps, CFG 3.0, hooks on, the FULL-SIZE
    UNet -- as the two B=4 windows SURVEY.md 8(d) allows (the reference materialises 24*B*4096^2 fp32 scores; the
    smoothing window of the reference is the DataLoader batch, so frame 4 has no predecessor).  One 8-frame clip from
    vface_b200.synth, frames [0,4) and [4,8); kept: the per-step latents of steps 1, 5 and 10 of each window."""
    from vface_b200 import synth
    from . import kernels as ok
    ref = rh.build_reference_unet()
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), seed=1))
    F, B, S = 8, 4, 10
    steps = ok.make_schedule(S)["ddim_timesteps"]
    clip = synth.synth_clip(F, steps=steps, flow_kind="smooth")
    out = {}
    for wi, lo in enumerate(range(0, F, B)):
        hi = lo + B
        sampler = rh.build_reference_sampler(ref)
        d = tempfile.mkdtemp()
        for t, v in clip["inversion"].items():
            torch.save(v[lo:hi].clone(), os.path.join(d, f"ddim_latents_{t}.pt"))
        with quiet():
            samples, inter = sampler.sample(
                S=S, batch_size=B, shape=(4, 64, 64), conditioning=clip["c"][lo:hi], target_conditioning=clip["target_cond"][lo:hi],
                inverse_results_dir=d, x_T=clip["x_T"][lo:hi], flow=clip["flow"][lo:hi - 1], unconditional_guidance_scale=3.0,
                unconditional_conditioning=clip["uc"][lo:hi], eta=0.0, verbose=False, log_every_t=1,
                test_model_kwargs=dict(inpaint_image=clip["inpaint_image"][lo:hi], inpaint_mask=clip["inpaint_mask"][lo:hi]))
        xs = torch.stack(inter["x_inter"][1:])
        out[f"x_inter_w{wi}"] = xs[[0, 4, 9]].numpy()
        print("window", wi, "done", flush=True)
    out["kept_steps"] = np.array([0, 4, 9])
    save("sampler_full_8f.npz", **out)


def golden_sampler_small_2way():
    """Row a4: DDIMSampler.p_sample_ddim (ddim_w_inv.py:564-617), the 2-way [uncond ; cond] step, on the reduced UNet.
    (1) called directly on an un-hooked UNet for three consecutive steps of a DDIM-10 schedule;
    (2) through sample(target_conditioning=None): ddim_sampling still registers the chunks=3 hooks (:289-305), which
        then split the 2B batch into thirds -- B = 3 so that 2B divides by 3 (any other B makes combine_fft_high_low's
        torch.cat fail in the reference); flow=None.  The mirror must reproduce that slicing, not 'fix' it."""
    from vface_b200 import synth
    from . import kernels as ok
    ref = rh.build_reference_unet(SMALL_UNET)
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), seed=1))
    S = 10
    out = {}
    # (1) direct calls, hooks never registered
    B = 2
    clip = synth.synth_clip(B)
    sampler = rh.build_reference_sampler(ref)
    with quiet():
        sampler.make_schedule(S, ddim_eta=0.0, verbose=False)
    x = clip["x_T"]
    time_range = np.flip(sampler.ddim_timesteps)
    xs, p0s = [], []
    kw = dict(test_model_kwargs=dict(inpaint_image=clip["inpaint_image"], inpaint_mask=clip["inpaint_mask"]))
    for i in range(3):
        ts = torch.full((B,), int(time_range[i]), dtype=torch.long)
        with quiet():
            x, p0 = sampler.p_sample_ddim(x, clip["c"], ts, index=S - 1 - i, unconditional_guidance_scale=3.0,
                                          unconditional_conditioning=clip["uc"], **kw)
        xs.append(x)
        p0s.append(p0)
    out["direct_x_prev"] = torch.stack(xs).numpy()
    out["direct_pred_x0"] = torch.stack(p0s).numpy()
    # scale 1.0: the single-branch path (:577-578)
    ts = torch.full((B,), int(time_range[0]), dtype=torch.long)
    with quiet():
        x1, _ = sampler.p_sample_ddim(clip["x_T"], clip["c"], ts, index=S - 1, unconditional_guidance_scale=1.0,
                                      unconditional_conditioning=clip["uc"], **kw)
    out["direct_x_prev_scale1"] = x1.numpy()
    # (2) through sample(), hooks registered by ddim_sampling, B = 3
    B = 3
    clip = synth.synth_clip(B)
    sampler = rh.build_reference_sampler(ref)
    with quiet():
        samples, inter = sampler.sample(
            S=S, batch_size=B, shape=(4, 64, 64), conditioning=clip["c"], target_conditioning=None,
            x_T=clip["x_T"], flow=None, unconditional_guidance_scale=3.0, unconditional_conditioning=clip["uc"],
            eta=0.0, verbose=False, log_every_t=1,
            test_model_kwargs=dict(inpaint_image=clip["inpaint_image"], inpaint_mask=clip["inpaint_mask"]))
    out["sample_x_inter"] = torch.stack(inter["x_inter"][1:]).numpy()
    save("sampler_small_2way.npz", **out)


def golden_sampler_small_eta():
    """Row a13: eta = 0.5 through the reference sampler on the reduced UNet.  noise_like (util.py:264-267) is called
    TWICE per step (ddim_w_inv.py:697 and :704) on the global CPU generator, seeded here with 123: the first draw of a
    step is the noise of x_prev, the second belongs to the discarded recon branch.  A mirror that draws once per step
    (or in another order) consumes a different sub-sequence and cannot match these latents."""
    from vface_b200 import synth
    from . import kernels as ok
    ref = rh.build_reference_unet(SMALL_UNET)
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), seed=1))
    B, S = 2, 5
    steps = ok.make_schedule(S)["ddim_timesteps"]
    clip = synth.synth_clip(B, steps=steps)
    sampler = rh.build_reference_sampler(ref)
    d = tempfile.mkdtemp()
    for t, v in clip["inversion"].items():
        torch.save(v, os.path.join(d, f"ddim_latents_{t}.pt"))
    torch.manual_seed(123)
    with quiet():
        samples, inter = sampler.sample(
            S=S, batch_size=B, shape=(4, 64, 64), conditioning=clip["c"], target_conditioning=clip["target_cond"],
            inverse_results_dir=d, x_T=clip["x_T"], flow=clip["flow"], unconditional_guidance_scale=3.0,
            unconditional_conditioning=clip["uc"], eta=0.5, verbose=False, log_every_t=1,
            test_model_kwargs=dict(inpaint_image=clip["inpaint_image"], inpaint_mask=clip["inpaint_mask"]))
    save("sampler_small_eta.npz", x_inter=torch.stack(inter["x_inter"][1:]).numpy(), seed=np.array(123), eta=np.array(0.5))


def golden_unet_full():
    """One forward of the full-size UNet (project_ffhq.yaml) on a 3-way batch of one frame."""
    from vface_b200 import synth
    ref = rh.build_reference_unet()
    ref.load_state_dict(synth.synth_state_dict(ref.state_dict(), seed=1))
    clip = synth.synth_clip(1)
    x = torch.cat([clip["x_T"], clip["inpaint_image"], clip["inpaint_mask"]], dim=1).repeat(3, 1, 1, 1)
    x[2, :4] = torch.randn(4, 64, 64, generator=torch.Generator().manual_seed(3))
    ctx = torch.cat([clip["uc"], clip["c"], clip["target_cond"]])
    with torch.no_grad():
        y = ref(x, torch.full((3,), 501, dtype=torch.long), context=ctx)
    save("unet_full.npz", out=y.numpy(), x_recon_lat=x[2, :4].numpy())


VAE_SMALL = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 4, 4],
                 num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def golden_vae_decoder():
    """Reference first-stage Decoder (model.py:462-570) + post_quant_conv (autoencoder.py:330-333) on a
    reduced ddconfig (ch 32: mid attention width 128), weights from vface_b200.synth (seed 3)."""
    from ldm.modules.diffusionmodules.model import Decoder
    from vface_b200 import synth
    with quiet():
        dec = Decoder(**VAE_SMALL).eval()
    pq = torch.nn.Conv2d(4, 4, 1)
    sd = synth.synth_state_dict({**{"decoder." + k: v for k, v in dec.state_dict().items()},
                                 **{"post_quant_conv." + k: v for k, v in pq.state_dict().items()}}, seed=3)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in sd.items() if k.startswith("post_quant_conv.")})
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        y = dec(pq(z / 0.18215))
    save("vae_decoder.npz", z=z.numpy(), out=y.numpy(), keys=np.array(sorted(sd.keys())))


def golden_vae_decoder_full():
    """The FULL first-stage decoder of the REFace config (project_ffhq.yaml first_stage_config.params.ddconfig: ch 128,
    ch_mult 1-2-4-4, mid AttnBlock = one head of width 512 over 64 x 64 = 4096 tokens) + post_quant_conv, unmodified
    reference on CPU, weights from vface_b200.synth (seed 3): one 64 x 64 latent -> 512 x 512 frame.  Kept: every 4th pixel
    of the frame (3 x 128 x 128) and the first 8 latent rows' worth of input is regenerated from the seed by the test."""
    from ldm.modules.diffusionmodules.model import Decoder
    from vface_b200 import synth
    full = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    with quiet():
        dec = Decoder(**full).eval()
    pq = torch.nn.Conv2d(4, 4, 1)
    sd = synth.synth_state_dict({**{"decoder." + k: v for k, v in dec.state_dict().items()},
                                 **{"post_quant_conv." + k: v for k, v in pq.state_dict().items()}}, seed=3)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in sd.items() if k.startswith("post_quant_conv.")})
    z = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(23))
    with torch.no_grad():
        y = dec(pq(z / 0.18215))
    save("vae_decoder_full.npz", out_strided=y[:, :, ::4, ::4].numpy(), out_mean=np.array(float(y.mean())),
         out_std=np.array(float(y.std())), seed=np.array(23))


def golden_vae_encoder():
    """Reference first-stage Encoder (model.py:368-459) + quant_conv (autoencoder.py:304, :322-326) on the reduced
    ddconfig, weights from vface_b200.synth (seed 5): image -> posterior moments; plus the posterior's mean / std."""
    from ldm.modules.diffusionmodules.model import Encoder
    from ldm.modules.distributions.distributions import DiagonalGaussianDistribution
    from vface_b200 import synth
    with quiet():
        enc = Encoder(**VAE_SMALL).eval()
    qc = torch.nn.Conv2d(8, 8, 1)
    sd = synth.synth_state_dict({**{"encoder." + k: v for k, v in enc.state_dict().items()},
                                 **{"quant_conv." + k: v for k, v in qc.state_dict().items()}}, seed=5)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    qc.load_state_dict({k[len("quant_conv."):]: v for k, v in sd.items() if k.startswith("quant_conv.")})
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(22))
    with torch.no_grad():
        moments = qc(enc(x))
        post = DiagonalGaussianDistribution(moments)
    save("vae_encoder.npz", x=x.numpy(), moments=moments.numpy(), mean=post.mean.numpy(), std=post.std.numpy(),
         keys=np.array(sorted(sd.keys())))


def block_projection_inputs(seed=77, batch=2, tokens=128, dim=320, ctx_dim=768):
    """Seeded inputs and parameters of golden_block_projection (regenerated by the tests: same torch build here and on
    the GPU box).  Parameters are rounded to bf16-representable values so that the bf16 CUDA path sees the same numbers."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0, shift=0.0: (torch.randn(*s, generator=g) * scale + shift).bfloat16().float()
    return dict(
        x=r(batch, tokens, dim, scale=1.3, shift=0.2), a=r(batch, tokens, dim), ctx=r(batch, 1, ctx_dim),
        g_in=r(batch, tokens, dim), t_out=r(batch, tokens, dim),
        n1_w=r(dim, scale=0.3, shift=1.0), n1_b=r(dim, scale=0.2),
        wq=r(dim, dim, scale=dim ** -0.5), wk=r(dim, dim, scale=dim ** -0.5), wv=r(dim, dim, scale=dim ** -0.5),
        wo1=r(dim, dim, scale=dim ** -0.5), bo1=r(dim, scale=0.3),
        wv2=r(dim, ctx_dim, scale=ctx_dim ** -0.5), wo2=r(dim, dim, scale=dim ** -0.5), bo2=r(dim, scale=0.3),
        w_in=r(dim, dim, scale=dim ** -0.5), b_in=r(dim, scale=0.3), w_out=r(dim, dim, scale=dim ** -0.5), b_out=r(dim, scale=0.3))


def golden_block_projection():
    """What the tcgen05 projection kernel (csrc/vf_gemm3.cu) replaces, computed by the reference's own modules
    (ldm/modules/attention.py: BasicTransformerBlock :224-243, CrossAttention :152-221, SpatialTransformer :246-288):
    norm1 + to_q/to_k/to_v; to_out(a) + x followed by attn2(norm2(.), single-token ctx) + .; proj_in; proj_out + x_in."""
    from ldm.modules.attention import BasicTransformerBlock, SpatialTransformer
    d = block_projection_inputs()
    dim = d["x"].shape[-1]
    with quiet():
        blk = BasicTransformerBlock(dim, 8, dim // 8, context_dim=d["ctx"].shape[-1]).eval()
        st = SpatialTransformer(dim, 8, dim // 8, depth=1, context_dim=d["ctx"].shape[-1]).eval()
    with torch.no_grad():
        blk.norm1.weight.copy_(d["n1_w"]); blk.norm1.bias.copy_(d["n1_b"])
        blk.attn1.to_q.weight.copy_(d["wq"]); blk.attn1.to_k.weight.copy_(d["wk"]); blk.attn1.to_v.weight.copy_(d["wv"])
        blk.attn1.to_out[0].weight.copy_(d["wo1"]); blk.attn1.to_out[0].bias.copy_(d["bo1"])
        blk.attn2.to_v.weight.copy_(d["wv2"]); blk.attn2.to_out[0].weight.copy_(d["wo2"]); blk.attn2.to_out[0].bias.copy_(d["bo2"])
        st.proj_in.weight.copy_(d["w_in"][:, :, None, None]); st.proj_in.bias.copy_(d["b_in"])
        st.proj_out.weight.copy_(d["w_out"][:, :, None, None]); st.proj_out.bias.copy_(d["b_out"])
        n1 = blk.norm1(d["x"])
        qkv = torch.cat([blk.attn1.to_q(n1), blk.attn1.to_k(n1), blk.attn1.to_v(n1)], dim=-1)
        x1 = blk.attn1.to_out(d["a"]) + d["x"]                                  # attention.py:239 with a = the attention output
        x2 = blk.attn2(blk.norm2(x1), context=d["ctx"]) + x1                    # attention.py:240-241, single-token context
        b, n, c = d["g_in"].shape
        hh, ww = 8, n // 8
        nchw = lambda t: t.reshape(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
        tok = lambda t: t.permute(0, 2, 3, 1).reshape(b, n, c)
        p_in = tok(st.proj_in(nchw(d["g_in"])))                                 # attention.py:279
        p_out = tok(st.proj_out(nchw(d["t_out"])) + nchw(d["x"]))               # attention.py:287-288
    save("block_projection.npz", qkv=qkv.numpy(), x2=x2.numpy(), proj_in=p_in.numpy(), proj_out=p_out.numpy())


def main():
    if not rh.available():
        sys.exit("reference not mounted; golden vectors can only be generated in the build container")
    rh.install()
    which = sys.argv[1:] or ["fsai", "warp", "attn_hooks", "schedule", "sampler_small", "unet_full", "sampler_full", "sampler_full_s10", "sampler_full_8f", "sampler_small_2way", "sampler_small_eta", "vae_decoder", "vae_decoder_full", "vae_encoder", "block_projection"]
    for w in which:
        globals()["golden_" + w]()


if __name__ == "__main__":
    main()
