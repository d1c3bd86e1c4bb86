"""First-stage decoder and the decoded-frame criterion of north_star.

  * Decoder + post_quant_conv against the UNMODIFIED reference Decoder (tests/golden/vae_decoder.npz,
    oracle/make_golden.py::golden_vae_decoder, reduced ddconfig).
  * "final decoded frames at PSNR >= 40 dB": the bf16 sampler's final latents and the reference's final
    latents (tests/golden/sampler_small.npz) are decoded by the same decoder and compared as images in
    [0, 1] (the script's clamp((x + 1) / 2, 0, 1), scripts/VFace_inference_batch.py:596-600).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
VAE_SMALL = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 4, 4],
                 num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def _decoder(dtype, ddconfig=VAE_SMALL, seed=3):
    from vface_b200 import synth
    from vface_b200.ldm.modules.diffusionmodules.model import AutoencoderKLDecoder
    m = AutoencoderKLDecoder(ddconfig)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=seed))
    return m.cuda().eval().to(dtype)


def psnr(a, b):
    img = lambda t: torch.clamp((t.double() + 1.0) / 2.0, 0.0, 1.0)
    mse = (img(a) - img(b)).pow(2).mean().item()
    return float("inf") if mse == 0 else 10.0 * np.log10(1.0 / mse)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
def test_decoder_vs_reference_golden(dtype, tol):
    gold = np.load(os.path.join(GOLD, "vae_decoder.npz"))
    m = _decoder(dtype)
    z = torch.from_numpy(gold["z"]).cuda()
    with torch.no_grad():
        y = m.decode((z / 0.18215).to(dtype)).float()
    want = torch.from_numpy(gold["out"]).cuda()
    assert tuple(y.shape) == tuple(want.shape)
    err = ((y - want).norm() / want.norm()).item()
    assert err < tol, err


def test_decode_first_stage_surface():
    """LatentDiffusion.decode_first_stage(z) = decoder(post_quant_conv(z / scale_factor)) (ddpm.py:1277-1284)."""
    from vface_b200 import synth
    from vface_b200.latent_diffusion import LatentDiffusion
    model = LatentDiffusion(unet_config=dict(model_channels=32, num_heads=2), first_stage_config=VAE_SMALL)
    fs = model.first_stage_model
    fs.load_state_dict(synth.synth_state_dict(fs.state_dict(), seed=3))
    model = model.cuda().eval()
    gold = np.load(os.path.join(GOLD, "vae_decoder.npz"))
    with torch.no_grad():
        y = model.decode_first_stage(torch.from_numpy(gold["z"]).cuda())
    want = torch.from_numpy(gold["out"]).cuda()
    assert ((y - want).norm() / want.norm()).item() < 2e-4


@pytest.mark.parametrize("kind", ["smooth", "integer"])
def test_decoded_frames_psnr_bf16_sampler_vs_reference(kind):
    """north_star: final decoded frames at PSNR >= 40 dB.  10 DDIM steps, hooks on, bf16 kernels; the
    reference's final latents come from the golden file; both go through the same fp32 decoder."""
    from oracle import kernels as ok
    from tests.test_pipeline_gpu import SMALL, build, run_sample
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_small.npz"))
    _, sampler, _ = build(SMALL, torch.bfloat16)
    S, B = 10, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind=kind)
    samples, _ = run_sample(sampler, clip, S, B, clip["inversion"])
    want = torch.from_numpy(gold[f"samples_{kind}"]).cuda()
    dec = _decoder(torch.float32, dict(VAE_SMALL, resolution=512))       # 64x64 latents -> 512x512 frames
    with torch.no_grad():
        a = dec.decode(samples.float() / 0.18215)
        b = dec.decode(want / 0.18215)
    assert tuple(a.shape) == (B, 3, 512, 512)
    # the comparison must not be vacuous: the decoded frames are not saturated / constant
    img = torch.clamp((b + 1) / 2, 0, 1)
    assert 0.02 < img.std().item() and 0.05 < img.mean().item() < 0.95
    p = psnr(a, b)
    assert p >= 40.0, p


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
def test_full_ddconfig_decoder_vs_reference_golden(dtype, tol):
    """The FULL REFace first-stage decoder (ddconfig of project_ffhq.yaml: ch 128, mid AttnBlock = one head of width 512 over
    4096 tokens) against the unmodified reference run on CPU (tests/golden/vae_decoder_full.npz: every 4th pixel of the
    512 x 512 frame).  In bf16 the AttnBlock runs on the repo's own tcgen05 attention kernel (wide-head form)."""
    from vface_b200 import ops
    gold = np.load(os.path.join(GOLD, "vae_decoder_full.npz"))
    m = _decoder(dtype, None)
    z = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(int(gold["seed"]))).cuda()
    n0 = ops.launch_count
    with torch.no_grad():
        y = m.decode((z / 0.18215).to(dtype)).float()
    assert tuple(y.shape) == (1, 3, 512, 512)
    want = torch.from_numpy(gold["out_strided"]).cuda()
    got = y[:, :, ::4, ::4]
    err = ((got - want).norm() / want.norm()).item()
    assert err < tol, err
    assert abs(y.mean().item() - float(gold["out_mean"])) < 5e-2 * max(1.0, float(gold["out_std"]))
    assert ops.launch_count > n0


def test_decoded_frames_psnr_full_size_bf16_sampler_vs_reference():
    """The same >= 40 dB bound on the FULL-SIZE path: the bf16 sampler over the 859.5 M-parameter UNet (DDIM 10 steps,
    BASELINE.json configs[0] schedule, hooks on) against the unmodified reference's final latents
    (tests/golden/sampler_full_s10.npz, last per-step latent = the samples), both decoded to 512x512 frames by the
    full REFace first-stage decoder (ddconfig of project_ffhq.yaml: ch 128, mid attention width 512; random-init weights)."""
    from oracle import kernels as ok
    from tests.test_pipeline_gpu import build, run_sample
    from vface_b200 import synth
    gold = np.load(os.path.join(GOLD, "sampler_full_s10.npz"))
    _, sampler, _ = build(None, torch.bfloat16)
    S, B = 10, 2
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="integer")
    samples, _ = run_sample(sampler, clip, S, B, clip["inversion"])
    del sampler
    torch.cuda.empty_cache()
    want = torch.from_numpy(gold["x_inter"][-1]).cuda()
    dec = _decoder(torch.float32, None)                                   # REFACE_DDCONFIG: 64x64 latents -> 512x512 frames
    with torch.no_grad():
        a = dec.decode(samples.float() / 0.18215)
        b = dec.decode(want / 0.18215)
    assert tuple(a.shape) == (B, 3, 512, 512)
    img = torch.clamp((b + 1) / 2, 0, 1)
    assert 0.02 < img.std().item() and 0.05 < img.mean().item() < 0.95
    p = psnr(a, b)
    assert p >= 40.0, p


# ---- encoder half (the save loop's encode -> decode round trip, scripts/VFace_inference_batch.py:456-459, :603-623) ----
def _autoencoder(dtype, seed_dec=3, seed_enc=5):
    from vface_b200 import synth
    from vface_b200.ldm.modules.diffusionmodules.model import AutoencoderKL
    m = AutoencoderKL(VAE_SMALL)
    sd = m.state_dict()
    dec = synth.synth_state_dict({k: v for k, v in sd.items() if k.startswith(("decoder.", "post_quant_conv."))}, seed=seed_dec)
    enc = synth.synth_state_dict({k: v for k, v in sd.items() if k.startswith(("encoder.", "quant_conv."))}, seed=seed_enc)
    m.load_state_dict({**dec, **enc})
    return m.cuda().eval().to(dtype)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
def test_encoder_vs_reference_golden(dtype, tol):
    """Encoder + quant_conv against the UNMODIFIED reference (tests/golden/vae_encoder.npz), same state-dict keys."""
    gold = np.load(os.path.join(GOLD, "vae_encoder.npz"))
    m = _autoencoder(dtype)
    enc_keys = sorted(k for k in m.state_dict() if k.startswith(("encoder.", "quant_conv.")))
    assert enc_keys == list(gold["keys"])
    x = torch.from_numpy(gold["x"]).cuda()
    with torch.no_grad():
        post = m.encode(x.to(dtype))
    want = torch.from_numpy(gold["moments"]).cuda()
    assert tuple(post.parameters.shape) == tuple(want.shape)
    err = ((post.parameters.float() - want).norm() / want.norm()).item()
    assert err < tol, err
    assert ((post.mean.float() - torch.from_numpy(gold["mean"]).cuda()).norm() / want.norm()).item() < tol
    if dtype == torch.float32:
        assert torch.allclose(post.std, torch.from_numpy(gold["std"]).cuda(), rtol=2e-3, atol=1e-6)
        assert torch.equal(post.mode(), post.mean)
        g = torch.Generator().manual_seed(1)
        z = post.sample(generator=g)
        assert z.shape == post.mean.shape and torch.isfinite(z).all()


def test_encode_decode_surface():
    """LatentDiffusion.encode_first_stage / get_first_stage_encoding / decode_first_stage keep the reference's contract
    (ddpm.py:782-791, :1277-1330): z = scale_factor * posterior sample, x' = decoder(post_quant_conv(z / scale_factor))."""
    from vface_b200.latent_diffusion import LatentDiffusion
    model = LatentDiffusion(unet_config=dict(model_channels=32, num_heads=2), first_stage_config=VAE_SMALL, first_stage_encoder=True)
    model.first_stage_model.load_state_dict(_autoencoder(torch.float32).state_dict())
    model = model.cuda().eval()
    gold = np.load(os.path.join(GOLD, "vae_encoder.npz"))
    x = torch.from_numpy(gold["x"]).cuda()
    with torch.no_grad():
        post = model.encode_first_stage(x)
        z = model.get_first_stage_encoding(post.mode())
        assert torch.allclose(z, 0.18215 * torch.from_numpy(gold["mean"]).cuda(), rtol=1e-3, atol=1e-5)
        y = model.decode_first_stage(z)
    assert tuple(y.shape) == (2, 3, 64, 64) and torch.isfinite(y).all()
    with pytest.raises(RuntimeError):
        LatentDiffusion(unet_config=dict(model_channels=32, num_heads=2), first_stage_config=VAE_SMALL).encode_first_stage(x)
