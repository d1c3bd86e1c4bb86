"""Full-size sampler at other frame counts (1 frame: no flow fields; 6 = the reference script's n_samples default;
5: odd): the bf16 path against the fp32 path (itself pinned to the reference at < 5e-6)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_pipeline_gpu import build, run_sample, rel_l2
from oracle import kernels as ok
from vface_b200 import synth
S = 2
for B in (1, 5, 6):
    outs = {}
    for dtype in (torch.float32, torch.bfloat16):
        _, sampler, _ = build(None, dtype)
        clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="smooth")
        samples, inter = run_sample(sampler, clip, S, B, clip["inversion"])
        outs[dtype] = [x.float().cpu() for x in inter["x_inter"][1:]]
        del sampler
        torch.cuda.empty_cache()
    errs = [rel_l2(a, b) for a, b in zip(outs[torch.bfloat16], outs[torch.float32])]
    print(f"B={B}: bf16 vs fp32 per-step rel L2 {[round(e, 5) for e in errs]}  finite={all(torch.isfinite(x).all().item() for x in outs[torch.bfloat16])}")
