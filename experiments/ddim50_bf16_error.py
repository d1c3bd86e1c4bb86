"""DDIM-50 (the BASELINE configuration), full-size UNet, 2 frames: per-step latent error of the bf16 path against the
fp32 path of the same kernels (which reproduces the reference to < 5e-6, tests/test_pipeline_gpu.py)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_pipeline_gpu import build, run_sample, rel_l2
from oracle import kernels as ok
from vface_b200 import synth
B = 2
for S in (50, 10):
    outs = {}
    for dtype in (torch.float32, torch.bfloat16):
        _, sampler, _ = build(None, dtype)
        clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="smooth")
        samples, inter = run_sample(sampler, clip, S, B, clip["inversion"])
        outs[dtype] = [x.float().cpu() for x in inter["x_inter"][1:]]
        del sampler
        torch.cuda.empty_cache()
    errs = np.array([rel_l2(a, b) for a, b in zip(outs[torch.bfloat16], outs[torch.float32])])
    print(f"S={S}: per-step rel L2 of bf16 vs fp32: first {errs[:5].round(5).tolist()} max {errs.max():.5f} (step {int(errs.argmax())}) mean {errs.mean():.5f} final {errs[-1]:.5f}")
    print("   all steps:", errs.round(4).tolist())
