"""Import the UNMODIFIED reference (Sanoojan/VFace, REFace/) on CPU.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference is mounted (the build container).  Nothing on
the GPU box may call this; callers must check `available()` first.

The shims below are harness-side only: no
reference file is edited or copied.  They stand in for packages that are absent
offline and that the hot path never exercises numerically:
  matplotlib / mpl_toolkits    imported by scripts/face_swap_utils.py:2-3 (debug plots)
  kornia.utils.create_meshgrid scripts/temporal_flow.py:14,43 (pixel grid; restated here)
  torchvision.io read/write    ldm/models/pnp_utils.py:10 (never called)
  raft_large                   scripts/temporal_flow.py:27 downloads weights at import
  omegaconf.listconfig         openaimodel.py:595 (type check only)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("VFACE_REFERENCE_ROOT", "/root/reference/REFace")

_installed = False


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "ldm", "models", "diffusion", "ddim_w_inv.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    # kornia 0.6 semantics: (1, H, W, 2) with [..., 0] = x, [..., 1] = y.
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    grid = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)
    return grid.permute(1, 0, 2).unsqueeze(0)


def install():
    """Put the reference on sys.path behind the shims.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = _mod("matplotlib")
            mpl.pyplot = _mod("matplotlib.pyplot")
    if "mpl_toolkits.mplot3d" not in sys.modules:
        try:
            import mpl_toolkits.mplot3d  # noqa: F401
        except Exception:
            _mod("mpl_toolkits")
            _mod("mpl_toolkits.mplot3d", Axes3D=object)
    try:
        import kornia  # noqa: F401
    except Exception:
        _mod("kornia", utils=_mod("kornia.utils", create_meshgrid=_create_meshgrid))
    import torchvision.io as tio
    if not hasattr(tio, "read_video"):
        tio.read_video = lambda *a, **k: None
    if not hasattr(tio, "write_video"):
        tio.write_video = lambda *a, **k: None
    import torchvision.models.optical_flow as tof
    tof.raft_large = lambda *a, **k: torch.nn.Identity()
    try:
        import omegaconf.listconfig  # noqa: F401
    except Exception:
        _mod("omegaconf")
        _mod("omegaconf.listconfig", ListConfig=type("ListConfig", (list,), {}))
    sys.path.insert(0, REF_ROOT)
    _installed = True


# UNet hyper-parameters of models/REFace/configs/project_ffhq.yaml:33-55
FULL_UNET = dict(image_size=32, in_channels=9, out_channels=4, model_channels=320,
                 attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                 num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
                 use_checkpoint=True, legacy=False)


class LatentDiffusionStub(torch.nn.Module):
    """What DDIMSampler and the hooks touch of ldm.models.diffusion.ddpm.LatentDiffusion
    (ddpm.py:255-277 schedule buffers, :1609 apply_model, :2245 DiffusionWrapper)."""

    def __init__(self, unet):
        super().__init__()
        from ldm.modules.diffusionmodules.util import make_beta_schedule
        self.model = torch.nn.Module()
        self.model.diffusion_model = unet
        self.model.conditioning_key = "crossattn"
        self.num_timesteps = 1000
        self.parameterization = "eps"
        betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
        alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", f32(betas))
        self.register_buffer("alphas_cumprod", f32(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", f32(alphas_cumprod_prev))

    @property
    def device(self):
        return self.betas.device

    def apply_model(self, x_noisy, t, cond):
        return self.model.diffusion_model(x_noisy, t, context=cond)


def build_reference_unet(unet_kwargs=None):
    install()
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    kw = dict(FULL_UNET)
    if unet_kwargs:
        kw.update(unet_kwargs)
    return UNetModel(**kw).eval()


def build_reference_sampler(unet):
    """Reference DDIMSampler on CPU; only register_buffer's forced .to('cuda')
    (ddim_w_inv.py:149-153) is neutralised."""
    install()
    from ldm.models.diffusion.ddim_w_inv import DDIMSampler

    class CpuDDIMSampler(DDIMSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    return CpuDDIMSampler(LatentDiffusionStub(unet))
