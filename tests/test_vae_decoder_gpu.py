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
