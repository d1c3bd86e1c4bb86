"""Convolutions of one 32-frame step (UNet batch 96) on cuDNN: time per distinct shape, cudnn.benchmark off/on.

    python benchmarks/bench_convs.py [--batch 96]
"""
import argparse, collections, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200.latent_diffusion import LatentDiffusion

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=96)
a = ap.parse_args()
unet = LatentDiffusion().model.diffusion_model
# record every conv call of one forward with its input shape
calls = collections.Counter()
orig = F.conv2d
def rec(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
    st = stride if isinstance(stride, int) else stride[0]
    pd = padding if isinstance(padding, int) else padding[0]
    calls[(tuple(x.shape[1:]), tuple(w.shape), st, pd)] += 1
    return orig(x, w, b, stride, padding, dilation, groups)
F.conv2d = rec
unet = unet.cuda().bfloat16()
with torch.no_grad():
    unet(torch.randn(1, 9, 64, 64, device="cuda"), torch.full((1,), 500, device="cuda"), torch.randn(1, 1, 768, device="cuda"))
F.conv2d = orig
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
tot = {False: 0.0, True: 0.0}
tot_fl = 0.0
print(f"{'input (c,h,w)':>18s} {'weight':>22s} s  calls   GFLOP   ms(off)  TF/s(off)   ms(bench)  TF/s(bench)")
for (xs, ws, st, pd), n in sorted(calls.items(), key=lambda kv: -kv[1]):
    if ws[2] == 1:
        continue       # 1x1 convolutions run as GEMMs in the product path
    x = torch.randn(a.batch, *xs, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(*ws, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    ho = (xs[1] + 2 * pd - ws[2]) // st + 1
    fl = 2.0 * a.batch * ho * ho * ws[0] * ws[1] * ws[2] * ws[3]
    res = {}
    for bench in (False, True):
        with torch.backends.cudnn.flags(enabled=True, benchmark=bench):
            res[bench] = timeit(lambda: F.conv2d(x, w, None, st, pd))
        tot[bench] += res[bench] * n
    tot_fl += fl * n
    print(f"{str(xs):>18s} {str(ws):>22s} {st}  {n:5d} {fl / 1e9:7.1f}  {res[False]:8.3f}  {fl / res[False] / 1e9:9.1f}  {res[True]:10.3f}  {fl / res[True] / 1e9:11.1f}")
print(f"total per step: {tot_fl / 1e12:.2f} TFLOP; {tot[False]:.2f} ms ({tot_fl / tot[False] / 1e9:.0f} TF/s) heuristic, {tot[True]:.2f} ms ({tot_fl / tot[True] / 1e9:.0f} TF/s) cudnn.benchmark")
