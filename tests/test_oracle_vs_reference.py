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
