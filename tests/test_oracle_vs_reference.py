"""The oracle against the LIVE reference, imported unmodified from /root/reference (build container
only; skipped elsewhere).  Complements test_oracle_vs_golden.py with randomised inputs."""
import numpy as np
import pytest
import torch

from oracle import kernels as ok
from oracle import port
from oracle import ref_harness as rh

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module", autouse=True)
def _ref():
    rh.install()


@pytest.mark.parametrize("d", [320, 640, 1280, 160])
def test_fsai_vs_live_reference(d):
    from scripts.face_swap_utils import combine_fft_high_low
    rng = np.random.default_rng(d)
    a = rng.standard_normal((2, 16, d)).astype(np.float32)
    b = rng.standard_normal((2, 16, d)).astype(np.float32)
    for r in (0.8, 0.3, 0.0, 1.0):
        want = combine_fft_high_low(torch.from_numpy(a), torch.from_numpy(b), split_ratio=r).numpy()
        assert np.abs(ok.fsai_blend(a, b, r) - want).max() < 5e-6


def test_warp_vs_live_reference():
    from scripts.temporal_flow import align_by_flow
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 6, 64, 64)).astype(np.float32)
    for scale in (2.0, 40.0):
        flow = (rng.standard_normal((3, 2, 64, 64)) * scale).astype(np.float32)
        want = align_by_flow(torch.from_numpy(x), flow=[torch.from_numpy(f)[None] for f in flow], alpha=0.8).numpy()
        tok = x.transpose(0, 2, 3, 1).reshape(4, 4096, 6)
        got = ok.flow_warp_blend(tok, flow, 0.8, 64, 64).reshape(4, 64, 64, 6).transpose(0, 3, 1, 2)
        assert np.abs(got - want).max() < 1e-6


def test_state_dict_keys_match_reference_full_size():
    from vface_b200.latent_diffusion import LatentDiffusion
    ref = rh.build_reference_unet()
    mine = LatentDiffusion().model.diffusion_model
    a = [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    b = [(k, tuple(v.shape)) for k, v in mine.state_dict().items()]
    assert a == b


def test_sampler_tables_match_reference():
    from vface_b200.latent_diffusion import LatentDiffusion
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    model = LatentDiffusion(unet=torch.nn.Identity())
    for S, eta in ((10, 0.0), (50, 0.0), (20, 0.7)):
        ref = rh.build_reference_sampler(torch.nn.Identity())
        ref.make_schedule(S, ddim_eta=eta, verbose=False)
        mine = DDIMSampler(model)
        mine.make_schedule(S, ddim_eta=eta, verbose=False)
        assert np.array_equal(mine.ddim_timesteps, ref.ddim_timesteps)
        for k, hk in (("ddim_alphas", "a_t"), ("ddim_alphas_prev", "a_prev"), ("ddim_sigmas", "sigma"),
                      ("ddim_sqrt_one_minus_alphas", "s1m")):
            r = np.asarray(torch.as_tensor(getattr(ref, k)).numpy(), dtype=np.float64).astype(np.float32)
            assert np.array_equal(mine._host_tables[hk], r), k


def test_install_rebinds_the_real_reference_tree():
    """vface_b200.install() against the REAL reference modules, in a fresh interpreter so the
    rebinding cannot leak into the other tests: every (module, name) pair is rebound, a module that imported a mirrored
    name earlier sees the new one, `instantiate_from_config` -- how scripts/VFace_inference_batch.py:118-135 builds the
    model from project_ffhq.yaml:33 -- returns the vface_b200 UNet, and its state-dict keys and shapes equal those of the
    reference UNet built before the rebinding (so last.ckpt loads unchanged)."""
    import os
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        from oracle import ref_harness as rh
        rh.install()                                           # shims + REFace/ on sys.path, nothing else
        import ldm.modules.attention, ldm.modules.diffusionmodules.openaimodel, ldm.modules.diffusionmodules.model
        import ldm.models.pnp_utils, ldm.models.diffusion.ddim_w_inv, scripts.face_swap_utils, scripts.temporal_flow
        from ldm.util import instantiate_from_config
        cfg = dict(target="ldm.modules.diffusionmodules.openaimodel.UNetModel",
                   params=dict(image_size=32, in_channels=9, out_channels=4, model_channels=32, attention_resolutions=[4, 2, 1],
                               num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=2, use_spatial_transformer=True,
                               transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False))
        ref_unet = instantiate_from_config(cfg)
        assert type(ref_unet).__module__ == "ldm.modules.diffusionmodules.openaimodel"
        ref_keys = {k: tuple(v.shape) for k, v in ref_unet.state_dict().items()}
        ref_sampler_cls = ldm.models.diffusion.ddim_w_inv.DDIMSampler

        import vface_b200
        done = vface_b200.install(strict=True)
        want = {(m, n) for m, names in vface_b200._PATCHED_NAMES.items() for n in names}
        assert want <= set(done), want - set(done)
        import importlib
        for m, n in want:
            obj = getattr(importlib.import_module(m), n)
            assert obj.__module__.startswith("vface_b200."), (m, n, obj.__module__)
        # late importers: pnp_utils did `from scripts.face_swap_utils import combine_fft_high_low` before install()
        assert ("ldm.models.pnp_utils", "combine_fft_high_low") in done
        assert ldm.models.pnp_utils.combine_fft_high_low.__module__.startswith("vface_b200.")
        assert ldm.modules.diffusionmodules.openaimodel.SpatialTransformer.__module__.startswith("vface_b200.")
        assert ldm.models.diffusion.ddim_w_inv.DDIMSampler is not ref_sampler_cls

        mine = instantiate_from_config(cfg)
        assert type(mine).__module__ == "vface_b200.ldm.modules.diffusionmodules.openaimodel", type(mine)
        got = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
        assert got == ref_keys
        mine.load_state_dict(ref_unet.state_dict(), strict=True)

        # the sampler the script constructs (VFace_inference_batch.py:19, :873) and the hook entry it reaches
        from ldm.models.diffusion.ddim_w_inv import DDIMSampler
        from ldm.models.pnp_utils import register_spa_attn_injection
        import inspect
        sig = inspect.signature(register_spa_attn_injection)
        assert list(sig.parameters)[:13] == ["model", "injection_schedule", "switch_on", "input_blocks", "output_blocks",
                                             "middle_block", "attn_component", "chunks", "flow", "block_indices", "fusion",
                                             "split_ratio_fft", "alpha"]
        stub = rh.LatentDiffusionStub(mine)
        s = DDIMSampler(stub)
        s.make_schedule(10, ddim_eta=0.0, verbose=False)
        s._register_hooks(None)
        mods = [m for n, m in mine.named_modules() if n.endswith("attn1")]
        assert len(mods) == 16 and all("forward" in m.__dict__ for m in mods)
        print("INSTALL_OK", len(done))
    """ % root)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "INSTALL_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
