import os, sys, re, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1]
dev = "cuda"
if "lib" in mode:
    from vface_b200 import _lib; _lib.load()
if "ltcall" in mode:
    from vface_b200 import ops
    x = torch.randn(512, 64, device=dev).bfloat16(); w = torch.randn(32, 64, device=dev).bfloat16(); r = torch.randn(512, 32, device=dev).bfloat16()
    ops.linear_residual(x, w, None, r); torch.cuda.synchronize(); print("  vf_linear_residual ran")
if "bias" in mode:
    x = torch.randn(6, 320, device=dev).bfloat16(); w = torch.randn(1280, 320, device=dev).bfloat16(); b = torch.randn(1280, device=dev).bfloat16()
    y = F.linear(x, w, b); torch.cuda.synchronize(); print("  F.linear with bias ok")
if "conv" in mode:
    x = torch.randn(6, 32, 64, 64, device=dev).bfloat16().contiguous(memory_format=torch.channels_last); w = torch.randn(32, 32, 3, 3, device=dev).bfloat16()
    y = F.conv2d(x, w, padding=1); torch.cuda.synchronize(); print("  conv ok")
x = torch.randn(6, 1, 768, device=dev).bfloat16(); w = torch.randn(32, 768, device=dev).bfloat16()
try:
    y = F.linear(x, w); torch.cuda.synchronize(); print("  F.linear no-bias ok")
except Exception as e:
    print("  F.linear no-bias FAILED:", str(e)[:70])
maps = open('/proc/self/maps').read()
print("  ", sorted(set(os.path.dirname(p) for p in re.findall(r'/\S*libcublas\S*', maps))))
