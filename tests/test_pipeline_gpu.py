"""GPU parity of the assembled hot path (UNet + hooks + sampler) against the reference.

Two checkers, both traceable to the reference:
  * tests/golden/*.npz      outputs of the UNMODIFIED reference run on CPU (oracle/make_golden.py)
  * oracle.port             the CPU restatement (pinned to the reference in test_oracle_vs_*.py)

Tolerances (BASELINE.json north_star): per-step latents within 1e-2 relative L2 (bf16); the fp32
path is held to 2e-4.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = dict(model_channels=32, num_heads=2)


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def build(unet_config=None, dtype=torch.float32, seed=1, **sampler_kw):
    from vface_b200 import synth
    from vface_b200.latent_diffusion import LatentDiffusion
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    model = LatentDiffusion(unet_config=unet_config)
    unet = model.model.diffusion_model
    sd = synth.synth_state_dict(unet.state_dict(), seed=seed)
    unet.load_state_dict(sd)
    model = model.cuda().eval()
    unet.to(dtype)
    return model, DDIMSampler(model, **sampler_kw), sd


def run_sample(sampler, clip, S, frames, inversion, flow=None, **kw):
    dev = "cuda"
    g = lambda t: t.to(dev)
    return sampler.sample(
        S=S, batch_size=frames, shape=(4, 64, 64), conditioning=g(clip["c"]), target_conditioning=g(clip["target_cond"]),
        inverse_results_dir=inversion, x_T=g(clip["x_T"]), flow=clip["flow"] if flow is None else flow,
        unconditional_guidance_scale=3.0, unconditional_conditioning=g(clip["uc"]), eta=0.0, verbose=False,
        log_every_t=1, test_model_kwargs=dict(inpaint_image=g(clip["inpaint_image"]), inpaint_mask=g(clip["inpaint_mask"])), **kw)


@pytest.mark.parametrize("kind", ["smooth", "integer"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 1e-2)])
def test_sampler_small_vs_reference_golden(kind, dtype, tol):
    """10 DDIM steps (BASELINE.json config 1), 2 frames, reduced UNet, hooks on (FSAI on 6 modules, flow warp on the 64x64 ones):
    every per-step latent against the reference's."""
    from oracle import kernels as ok
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_small.npz"))
    _, sampler, _ = build(SMALL, dtype)
    S, B = 10, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind=kind)
    samples, inter = run_sample(sampler, clip, S, B, clip["inversion"])
    want = gold[f"x_inter_{kind}"]
    for i in range(S):
        err = rel_l2(inter["x_inter"][1 + i], want[i])
        assert err < tol, (i, err)
    assert rel_l2(samples, gold[f"samples_{kind}"]) < tol
    # pred_x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t) is not a per-step latent: at the first steps (t = 901:
    # sqrt(a_t) = 0.08) it divides the eps error by sqrt(a_t), so in bf16 it is held to 2x the latent bound
    # over the whole trajectory and to the latent bound itself over the second half (a_t >= 0.27).
    px0, want0 = torch.stack(inter["pred_x0"][1:]), gold[f"pred_x0_{kind}"]
    assert rel_l2(px0, want0) < (tol if dtype == torch.float32 else 2 * tol)
    assert rel_l2(px0[S // 2:], want0[S // 2:]) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 1e-2)])
def test_sampler_full_size_vs_reference_golden(dtype, tol):
    """The FULL-SIZE UNet (859.5 M parameters) through DDIMSampler.sample: 2 frames, 5 DDIM steps, CFG 3.0, hooks on
    (FSAI on six modules, flow warp on the two 64x64 ones): every per-step latent against the unmodified reference
    run on CPU (tests/golden/sampler_full.npz, oracle/make_golden.py::golden_sampler_full)."""
    from oracle import kernels as ok
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_full.npz"))
    _, sampler, _ = build(None, dtype)
    S, B = 5, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="smooth")
    samples, inter = run_sample(sampler, clip, S, B, clip["inversion"])
    want = gold["x_inter"]
    assert float(np.abs(want[-1] - want[0]).mean()) > 0.05          # the trajectory moves: not a vacuous comparison
    for i in range(S):
        err = rel_l2(inter["x_inter"][1 + i], want[i])
        assert err < tol, (i, err)
    assert rel_l2(samples, gold["samples"]) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 1e-2)])
def test_sampler_full_size_ddim10_vs_reference_golden(dtype, tol):
    """BASELINE.json configs[0] (DDIM 10 steps, CFG 3.0, hooks on) on the full-size UNet, 2 frames, integer-valued flow:
    all ten per-step latents against the unmodified reference run on CPU (tests/golden/sampler_full_s10.npz)."""
    from oracle import kernels as ok
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_full_s10.npz"))
    _, sampler, _ = build(None, dtype)
    S, B = 10, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="integer")
    _, inter = run_sample(sampler, clip, S, B, clip["inversion"])
    want = gold["x_inter"]
    assert want.shape[0] == S
    errs = [rel_l2(inter["x_inter"][1 + i], want[i]) for i in range(S)]
    assert max(errs) < tol, errs


def test_ddim50_full_size_bf16_per_step_error():
    """BASELINE.json's configuration (DDIM-50, CFG 3.0, full-size UNet, hooks on): every per-step latent of the bf16
    path within 1e-2 relative L2 of the fp32 path of the same kernels -- which reproduces the unmodified reference to
    < 5e-6 on this UNet (test_sampler_full_size_vs_reference_golden[float32]); a 50-step CPU run of the reference
    itself takes 25 minutes, so the S=50 trajectory is pinned transitively.  Measured: 0.0020 at the first step,
    0.0053 at the last (coarser schedules take larger steps: 0.0075 at S=10, 0.0091 at S=5)."""
    from oracle import kernels as ok
    from vface_b200 import synth
    S, B = 50, 2
    traj = {}
    for dtype in (torch.float32, torch.bfloat16):
        _, sampler, _ = build(None, dtype)
        clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="smooth")
        _, inter = run_sample(sampler, clip, S, B, clip["inversion"])
        traj[dtype] = [x.float().cpu() for x in inter["x_inter"][1:]]
        del sampler
        torch.cuda.empty_cache()
    assert len(traj[torch.float32]) == S
    errs = [rel_l2(a, b) for a, b in zip(traj[torch.bfloat16], traj[torch.float32])]
    assert max(errs) < 1e-2, (int(np.argmax(errs)), max(errs))
    assert float((traj[torch.float32][-1] - traj[torch.float32][0]).abs().mean()) > 0.05      # the trajectory moves


def test_inversion_dir_and_dict_agree(tmp_path):
    """The reference's on-disk format (ddim_latents_{t}.pt per step) and the in-memory hand-off give
    identical samples."""
    from oracle import kernels as ok
    from vface_b200 import synth
    _, sampler, _ = build(SMALL, torch.float32)
    S, B = 10, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"])
    for t, v in clip["inversion"].items():
        torch.save(v, tmp_path / f"ddim_latents_{t}.pt")
    a, _ = run_sample(sampler, clip, S, B, clip["inversion"])
    b, _ = run_sample(sampler, clip, S, B, str(tmp_path))
    assert torch.equal(a, b)


def test_ddim_invert_small_vs_reference_golden(tmp_path):
    from oracle import kernels as ok
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_small.npz"))
    _, sampler, _ = build(SMALL, torch.float32)
    S, B = 10, 2
    clip = synth.synth_clip(2 * B)
    g = lambda t: t.cuda()
    xT, inter = sampler.ddim_invert(x=g(clip["x_T"]), cond=g(clip["c"]), S=S, shape=(4, 64, 64), eta=0.0,
                                    inverse_dir=str(tmp_path), batch_size=B,
                                    test_model_kwargs=dict(inpaint_image=g(clip["inpaint_image"]), inpaint_mask=g(clip["inpaint_mask"])))
    assert rel_l2(xT, gold["invert_xT"]) < 2e-4
    for t in ok.make_schedule(S)["ddim_timesteps"]:
        disk = torch.load(tmp_path / f"ddim_latents_{int(t)}.pt")
        assert rel_l2(disk, gold[f"invert_saved_{int(t)}"]) < 2e-4
        assert torch.equal(disk.cuda(), sampler.last_inversion[int(t)])


def test_elided_recon_branch_matches_three_branch():
    """the recon branch never reaches the output; skipping it changes nothing."""
    from oracle import kernels as ok
    from vface_b200 import synth
    S, B = 10, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"])
    _, s3, _ = build(SMALL, torch.float32)
    _, s2, _ = build(SMALL, torch.float32, elide_dead_recon=True)
    a, _ = run_sample(s3, clip, S, B, clip["inversion"])
    b, _ = run_sample(s2, clip, S, B, clip["inversion"])
    assert rel_l2(b, a) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
def test_unet_full_size_vs_reference_golden(dtype, tol):
    """One forward of the full 859.5 M-parameter UNet (project_ffhq.yaml) against the reference's output.
    The output here is epsilon, not a latent: the 1e-2 bound of north_star applies to per-step latents
    (tested below and in test_sampler_small_*), into which epsilon enters with a coefficient < 1 at
    S >= 10; the bf16 budget for epsilon itself (weights and activations rounded to bf16 through
    ~60 layers) is set at 2e-2 relative L2, measured 1.4e-2."""
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "unet_full.npz"))
    model, _, _ = build(None, dtype)
    clip = synth.synth_clip(1)
    x = torch.cat([clip["x_T"], clip["inpaint_image"], clip["inpaint_mask"]], dim=1).repeat(3, 1, 1, 1)
    x[2, :4] = torch.from_numpy(gold["x_recon_lat"])
    ctx = torch.cat([clip["uc"], clip["c"], clip["target_cond"]])
    with torch.no_grad():
        y = model.apply_model(x.cuda(), torch.full((3,), 501, dtype=torch.long, device="cuda"), ctx.cuda())
    assert y.dtype == torch.float32 and tuple(y.shape) == (3, 4, 64, 64)
    assert rel_l2(y, gold["out"]) < tol


def test_full_size_step_vs_oracle_port():
    """One full-size denoising step with hooks and flow (1 frame cannot warp, so 2 frames) against the
    CPU port, bf16 kernels: per-step latents within 1e-2 relative L2."""
    from oracle import kernels as ok, port
    from vface_b200 import synth
    S, B = 10, 2
    steps = ok.make_schedule(S)["ddim_timesteps"]
    clip = synth.synth_clip(B, steps=steps)
    model, sampler, sd = build(None, torch.bfloat16)
    want, xs, _ = port.sample(sd, 8, S, clip["x_T"], clip["c"], clip["target_cond"], clip["uc"], clip["inpaint_image"],
                              clip["inpaint_mask"], clip["inversion"], clip["flow"], return_all=True, max_steps=1)
    sampler.make_schedule(S, verbose=False)
    sampler._register_hooks(clip["flow"])
    g = lambda t: t.cuda()
    step = int(steps[-1])
    x_prev, pred_x0 = sampler.p_sample_ddim_with_inverse(
        g(clip["x_T"]), g(clip["c"]), torch.full((B,), step, device="cuda", dtype=torch.long), index=S - 1,
        target_conditioning=g(clip["target_cond"]), inverse_results_dir={k: g(v) for k, v in clip["inversion"].items()},
        unconditional_guidance_scale=3.0, unconditional_conditioning=g(clip["uc"]), flow=clip["flow"],
        test_model_kwargs=dict(inpaint_image=g(clip["inpaint_image"]), inpaint_mask=g(clip["inpaint_mask"])))
    assert rel_l2(x_prev, xs[0]) < 1e-2


# ---- rows a4 / a13 / (b): the 2-way step, the RNG draws, the script's autocast ----------------------------
def test_p_sample_ddim_direct_vs_reference_golden():
    """Row a4: DDIMSampler.p_sample_ddim (ddim_w_inv.py:564-617), the 2-way [uncond ; cond] step, called directly for
    three consecutive steps on an un-hooked reduced UNet, against the unmodified reference; plus its single-branch path
    (unconditional_guidance_scale == 1, :577-578)."""
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_small_2way.npz"))
    _, sampler, _ = build(SMALL, torch.float32)
    S, B = 10, 2
    clip = synth.synth_clip(B)
    sampler.make_schedule(S, ddim_eta=0.0, verbose=False)
    g = lambda t: t.cuda()
    kw = dict(test_model_kwargs=dict(inpaint_image=g(clip["inpaint_image"]), inpaint_mask=g(clip["inpaint_mask"])))
    time_range = np.flip(sampler.ddim_timesteps)
    x = g(clip["x_T"])
    for i in range(3):
        ts = torch.full((B,), int(time_range[i]), device="cuda", dtype=torch.long)
        x, p0 = sampler.p_sample_ddim(x, g(clip["c"]), ts, index=S - 1 - i, unconditional_guidance_scale=3.0,
                                      unconditional_conditioning=g(clip["uc"]), **kw)
        assert rel_l2(x, gold["direct_x_prev"][i]) < 2e-4, i
        assert rel_l2(p0, gold["direct_pred_x0"][i]) < 2e-4, i
    ts = torch.full((B,), int(time_range[0]), device="cuda", dtype=torch.long)
    x1, _ = sampler.p_sample_ddim(g(clip["x_T"]), g(clip["c"]), ts, index=S - 1, unconditional_guidance_scale=1.0,
                                  unconditional_conditioning=g(clip["uc"]), **kw)
    assert rel_l2(x1, gold["direct_x_prev_scale1"]) < 2e-4


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 1e-2)])
def test_sample_without_target_conditioning_vs_reference_golden(dtype, tol):
    """Row a4 through the public entry: sample(target_conditioning=None) dispatches to p_sample_ddim (:337-344) while
    ddim_sampling has still registered the chunks=3 hooks (:289-305), which then cut the 2B batch into thirds
    (pnp_utils.py:97-101).  B = 3 (the only kind of B the reference itself survives); the mirror reproduces the
    reference's slicing, it does not 'fix' it."""
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_small_2way.npz"))
    _, sampler, _ = build(SMALL, dtype)
    S, B = 10, 3
    clip = synth.synth_clip(B)
    g = lambda t: t.cuda()
    samples, inter = sampler.sample(
        S=S, batch_size=B, shape=(4, 64, 64), conditioning=g(clip["c"]), target_conditioning=None, x_T=g(clip["x_T"]),
        flow=None, unconditional_guidance_scale=3.0, unconditional_conditioning=g(clip["uc"]), eta=0.0, verbose=False,
        log_every_t=1, test_model_kwargs=dict(inpaint_image=g(clip["inpaint_image"]), inpaint_mask=g(clip["inpaint_mask"])))
    want = gold["sample_x_inter"]
    errs = [rel_l2(inter["x_inter"][1 + i], want[i]) for i in range(S)]
    assert max(errs) < tol, errs


def test_eta_noise_two_draws_per_step_vs_reference_golden(monkeypatch):
    """Row a13: the reference calls noise_like TWICE per step (ddim_w_inv.py:697 for x_prev, :704 for the discarded recon
    branch).  tests/golden/sampler_small_eta.npz is the unmodified reference at eta = 0.5 with the global CPU generator
    seeded 123; here noise_like is replaced by draws from a CPU generator with that seed, so the trajectories agree only
    if the mirror consumes the same sub-sequence: two draws per step, the FIRST one used."""
    from oracle import kernels as ok
    from vface_b200 import synth
    from vface_b200.ldm.models.diffusion import ddim_w_inv as mod
    gold = np.load(os.path.join(GOLD, "sampler_small_eta.npz"))
    gen = torch.Generator(device="cpu").manual_seed(int(gold["seed"]))
    calls = []

    def cpu_noise_like(shape, device, repeat=False):
        assert not repeat
        calls.append(tuple(shape))
        return torch.randn(shape, generator=gen).to(device)

    monkeypatch.setattr(mod, "noise_like", cpu_noise_like)
    _, sampler, _ = build(SMALL, torch.float32)
    S, B = 5, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"])
    g = lambda t: t.cuda()
    samples, inter = sampler.sample(
        S=S, batch_size=B, shape=(4, 64, 64), conditioning=g(clip["c"]), target_conditioning=g(clip["target_cond"]),
        inverse_results_dir=clip["inversion"], x_T=g(clip["x_T"]), flow=clip["flow"], unconditional_guidance_scale=3.0,
        unconditional_conditioning=g(clip["uc"]), eta=float(gold["eta"]), verbose=False, log_every_t=1,
        test_model_kwargs=dict(inpaint_image=g(clip["inpaint_image"]), inpaint_mask=g(clip["inpaint_mask"])))
    assert calls == [(B, 4, 64, 64)] * (2 * S)
    errs = [rel_l2(inter["x_inter"][1 + i], gold["x_inter"][i]) for i in range(S)]
    assert max(errs) < 2e-4, errs


def test_noise_draws_advance_cuda_generator_twice_per_step():
    """Row a13 on the real generator: after one p_sample_ddim_with_inverse step (eta > 0) the CUDA generator is where two
    torch.randn(B, 4, 64, 64) draws leave it, and x_prev carries the FIRST draw:
    x_prev - x_prev(eta-free part) = sigma_t * noise_1 (ddim_w_inv.py:696-700)."""
    from oracle import kernels as ok
    from vface_b200 import synth
    _, sampler, _ = build(SMALL, torch.float32)
    S, B = 5, 2
    steps = ok.make_schedule(S)["ddim_timesteps"]
    clip = synth.synth_clip(B, steps=steps)
    g = lambda t: t.cuda()
    sampler.make_schedule(S, ddim_eta=0.7, verbose=False)
    sampler._register_hooks(clip["flow"])
    step = int(steps[-1])
    args = (g(clip["x_T"]), g(clip["c"]), torch.full((B,), step, device="cuda", dtype=torch.long))
    kw = dict(index=S - 1, target_conditioning=g(clip["target_cond"]), inverse_results_dir={k: g(v) for k, v in clip["inversion"].items()},
              unconditional_guidance_scale=3.0, unconditional_conditioning=g(clip["uc"]), flow=clip["flow"],
              test_model_kwargs=dict(inpaint_image=g(clip["inpaint_image"]), inpaint_mask=g(clip["inpaint_mask"])))
    torch.manual_seed(99)
    x_prev, pred_x0 = sampler.p_sample_ddim_with_inverse(*args, **kw)
    state_after_step = torch.cuda.get_rng_state()
    torch.manual_seed(99)
    n1 = torch.randn(B, 4, 64, 64, device="cuda")
    _n2 = torch.randn(B, 4, 64, 64, device="cuda")
    assert torch.equal(torch.cuda.get_rng_state(), state_after_step)
    tb = sampler._host_tables
    a_t, a_prev, sigma, s1m = (float(tb[k][S - 1]) for k in ("a_t", "a_prev", "sigma", "s1m"))
    assert sigma > 0
    e = (g(clip["x_T"]) - pred_x0 * a_t ** 0.5) / s1m
    det = a_prev ** 0.5 * pred_x0 + (1.0 - a_prev - sigma ** 2) ** 0.5 * e
    assert rel_l2(x_prev - det, sigma * n1) < 1e-3


@pytest.mark.parametrize("start_dtype", [torch.float32, torch.bfloat16])
def test_sampler_under_script_default_autocast(start_dtype):
    """The drop-in target samples inside `with autocast("cuda")` by default (scripts/VFace_inference_batch.py:400,
    :407-409; fp16).  With fp32 parameters the mirror adopts bf16 once, with bf16 parameters nothing changes; either way
    autocast is off inside the mirrored forward (no float16 reaches a kernel) and every per-step latent stays within the
    bf16 bound of the unmodified reference's."""
    import warnings
    from oracle import kernels as ok
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_small.npz"))
    model, sampler, _ = build(SMALL, start_dtype)
    S, B = 10, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="smooth")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with torch.autocast("cuda"):
            samples, inter = run_sample(sampler, clip, S, B, clip["inversion"])
    assert model.model.diffusion_model.dtype == torch.bfloat16
    assert (start_dtype == torch.float32) == any("autocast" in str(x.message) for x in w)
    assert samples.dtype == torch.float32
    errs = [rel_l2(inter["x_inter"][1 + i], gold["x_inter_smooth"][i]) for i in range(S)]
    assert max(errs) < 1e-2, errs


def test_hooked_attention_under_autocast_direct_call():
    """A foreign caller invoking the patched attn1.forward inside an autocast region (no UNetModel.forward around it):
    the closure switches autocast off itself."""
    from vface_b200.ldm.modules.attention import CrossAttention
    from vface_b200.ldm.models.pnp_utils import register_spa_attn_injection
    torch.manual_seed(0)
    attn = CrossAttention(query_dim=320, heads=8, dim_head=40).cuda().to(torch.bfloat16)
    blk = torch.nn.Module(); blk.attn1 = attn
    unet = torch.nn.Module()
    unet.input_blocks = torch.nn.ModuleList([blk]); unet.middle_block = torch.nn.ModuleList([]); unet.output_blocks = torch.nn.ModuleList([])

    class H:
        pass
    h = H(); h.model = H(); h.model.model = H(); h.model.model.diffusion_model = unet
    register_spa_attn_injection(h, 1, switch_on=True, input_blocks=True, output_blocks=False, middle_block=False,
                                attn_component="attn1", chunks=3, fusion="fft", split_ratio_fft=0.8)
    x = torch.randn(3, 256, 320, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        want = attn.forward(x.clone())
        with torch.autocast("cuda"):
            got = attn.forward(x.clone())
    assert got.dtype == torch.bfloat16 and torch.equal(got, want)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 1e-2)])
def test_config0_eight_frames_vs_reference_golden(dtype, tol):
    """BASELINE.json configs[0] at its full frame count: 8-frame clip, DDIM 10 steps, CFG 3.0, hooks on, full-size UNet,
    as the two B=4 windows SURVEY.md 8(d) allows (tests/golden/sampler_full_8f.npz: the unmodified reference on CPU,
    per-step latents of steps 1, 5 and 10 of both windows).  The smoothing window of the reference is the batch, so each
    window is one sample() call: frames [0,4) and [4,8) of one 8-frame synthetic clip."""
    from oracle import kernels as ok
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_full_8f.npz"))
    _, sampler, _ = build(None, dtype)
    F, B, S = 8, 4, 10
    clip = synth.synth_clip(F, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="smooth")
    kept = [int(i) for i in gold["kept_steps"]]
    for wi, lo in enumerate(range(0, F, B)):
        hi = lo + B
        win = {k: v[lo:hi] for k, v in clip.items() if isinstance(v, torch.Tensor)}
        win["flow"] = clip["flow"][lo:hi - 1]
        inv = {t: v[lo:hi] for t, v in clip["inversion"].items()}
        _, inter = run_sample(sampler, win, S, B, inv)
        want = gold[f"x_inter_w{wi}"]
        errs = [rel_l2(inter["x_inter"][1 + s], want[j]) for j, s in enumerate(kept)]
        assert max(errs) < tol, (wi, errs)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cuda_graph_steps_match_eager(dtype):
    """Extension: DDIMSampler(cuda_graphs=True) captures one CUDA graph per schedule position on the first
    sample() call and replays them on the next; both calls, on two different clips, reproduce the eager sampler bit for bit
    (same kernels, same order)."""
    from oracle import kernels as ok
    from vface_b200 import synth
    S, B = 10, 2
    steps = ok.make_schedule(S)["ddim_timesteps"]
    _, eager, _ = build(SMALL, dtype)
    _, graphed, _ = build(SMALL, dtype, cuda_graphs=True)
    for seed in (7, 31):
        clip = synth.synth_clip(B, seed=seed, steps=steps)
        a, ia = run_sample(eager, clip, S, B, clip["inversion"])
        b, ib = run_sample(graphed, clip, S, B, clip["inversion"])
        assert torch.equal(a, b), seed
        assert all(torch.equal(x, y) for x, y in zip(ia["x_inter"], ib["x_inter"]))
    assert len(graphed._graphs) == S            # captured once, replayed for the second clip
