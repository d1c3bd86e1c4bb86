import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_pipeline_gpu import build, run_sample, rel_l2, GOLD
from oracle import kernels as ok
from vface_b200 import synth
gold = np.load(os.path.join(GOLD, "sampler_full.npz"))
S, B = 5, 2
for dtype in (torch.float32, torch.bfloat16):
    _, sampler, _ = build(None, dtype)
    clip = synth.synth_clip(B, steps=ok.make_schedule(S)["ddim_timesteps"], flow_kind="smooth")
    samples, inter = run_sample(sampler, clip, S, B, clip["inversion"])
    print(dtype, [round(rel_l2(inter["x_inter"][1 + i], gold["x_inter"][i]), 5) for i in range(S)], "pred_x0", [round(rel_l2(inter["pred_x0"][1 + i], gold["pred_x0"][i]), 5) for i in range(S)])
    print("  norms", [float(np.linalg.norm(gold["x_inter"][i])) for i in range(S)])
# bf16 with fp32 attention / hooks only? isolate: eps error of one UNet call at the first step
