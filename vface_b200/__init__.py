"""vface_b200 -- B200-native implementation of the VFace denoising hot path.

Hand-written sm_100a kernels (csrc/, C-ABI in include/vface_b200.h) behind the reference's own
Python call surface (`vface_b200.ldm...` mirrors REFace/ldm/...).  `install()` makes the mirror
importable under the reference's module names so scripts/VFace_inference_batch.py runs as a drop-in
(see INTEGRATION.md).
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

_MIRRORED = {
    "ldm.modules.diffusionmodules.util": "vface_b200.ldm.modules.diffusionmodules.util",
    "ldm.modules.diffusionmodules.openaimodel": "vface_b200.ldm.modules.diffusionmodules.openaimodel",
    "ldm.modules.attention": "vface_b200.ldm.modules.attention",
    "ldm.models.pnp_utils": "vface_b200.ldm.models.pnp_utils",
    "ldm.models.diffusion.ddim_w_inv": "vface_b200.ldm.models.diffusion.ddim_w_inv",
}

_PATCHED_NAMES = {
    "ldm.modules.diffusionmodules.openaimodel": ["UNetModel", "ResBlock", "Upsample", "Downsample", "TimestepEmbedSequential"],
    "ldm.modules.attention": ["CrossAttention", "BasicTransformerBlock", "SpatialTransformer", "FeedForward", "GEGLU"],
    "ldm.models.pnp_utils": ["register_spa_attn_injection", "find_all_modules_by_name"],
    "ldm.models.diffusion.ddim_w_inv": ["DDIMSampler", "load_ddim_latents_at_t"],
    "ldm.modules.diffusionmodules.model": ["Decoder", "Encoder", "ResnetBlock", "AttnBlock", "Upsample", "Downsample"],
    "scripts.face_swap_utils": ["combine_fft_high_low"],
    "scripts.temporal_flow": ["align_by_flow", "warp_image"],
}

_SOURCES = dict(_MIRRORED, **{"ldm.modules.diffusionmodules.model": "vface_b200.ldm.modules.diffusionmodules.model",
                              "scripts.face_swap_utils": "vface_b200.scripts.face_swap_utils",
                              "scripts.temporal_flow": "vface_b200.scripts.temporal_flow"})

# Modules of the reference that bind mirrored names with `from X import name` at THEIR import time: if one of them is
# already loaded when install() runs, its copy of the name is rebound too (module -> names it may hold).
_IMPORTERS = {
    "ldm.models.autoencoder": ["Encoder", "Decoder"],                                   # autoencoder.py:8
    "ldm.models.diffusion.ddim_w_inv": ["register_spa_attn_injection"],                 # ddim_w_inv.py:17 (replaced wholesale anyway)
    "ldm.models.pnp_utils": ["combine_fft_high_low", "align_by_flow"],
    "ldm.modules.diffusionmodules.openaimodel": ["SpatialTransformer"],                 # openaimodel.py:20
}


def install(strict: bool = False):
    """Drop-in: rebind the hot-path classes/functions of the reference's modules to vface_b200's.

    Call after the reference tree (REFace/) is on sys.path and BEFORE the model is instantiated from
    its config.  Modules of the reference that cannot be imported are skipped unless `strict`.
    Returns the list of (module, name) pairs that were rebound.
    """
    done = []
    for ref_name, names in _PATCHED_NAMES.items():
        mine = importlib.import_module(_SOURCES[ref_name])
        try:
            ref_mod = importlib.import_module(ref_name)
        except Exception:
            if strict:
                raise
            continue
        for n in names:
            setattr(ref_mod, n, getattr(mine, n))
            done.append((ref_name, n))
    rebound = {n: obj for (m, n) in done for obj in [getattr(sys.modules[m], n)]}
    for mod_name, names in _IMPORTERS.items():
        mod = sys.modules.get(mod_name)
        if mod is None or mod.__name__.startswith("vface_b200"):
            continue
        for n in names:
            if n in rebound and hasattr(mod, n) and getattr(mod, n) is not rebound[n]:
                setattr(mod, n, rebound[n])
                done.append((mod_name, n))
    return done


def alias_modules():
    """Alternative to install() when the reference tree is absent: expose the mirror under the
    reference's module names (import ldm.models.diffusion.ddim_w_inv -> vface_b200's)."""
    for ref_name, mine in _MIRRORED.items():
        sys.modules.setdefault(ref_name, importlib.import_module(mine))
