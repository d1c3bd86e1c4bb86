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
