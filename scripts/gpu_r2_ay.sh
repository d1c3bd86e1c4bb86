#!/bin/bash
# round-2 GPU call AY: gn_ws_kernel with bulk stores out of shared memory for single-source slabs (VF_GN_WS=2)
mkdir -p gpurun_out
VF_GN_WS=2 timeout 300 python - > gpurun_out/r2ay_check.log 2>&1 <<'PY'
import torch, torch.nn.functional as F
from vface_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
for n, hw, c1, c2, add in ((96, 4096, 320, 0, True), (96, 1024, 640, 320, False), (24, 256, 1280, 0, True), (96, 1024, 640, 0, False)):
    c = c1 + c2
    x1 = (torch.randn(n, hw, c1, generator=g) * 1.5 + 0.3).bfloat16().to(dev)
    x2 = torch.randn(n, hw, c2, generator=g).bfloat16().to(dev) if c2 else None
    w, b = torch.randn(c, generator=g).bfloat16().to(dev), torch.randn(c, generator=g).bfloat16().to(dev)
    a = torch.randn(n, c, generator=g).bfloat16().to(dev) if add else None
    got = ops.group_norm_nhwc(x1, w, b, 1e-5, 32, silu=True, add_nc=a, x2=x2)
    again = ops.group_norm_nhwc(x1, w, b, 1e-5, 32, silu=True, add_nc=a, x2=x2)
    assert torch.equal(got, again)
    xf = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)
    if a is not None:
        xf = xf + a.float()[:, None, :]
    want = F.silu(F.group_norm(xf.permute(0, 2, 1), 32, w.float(), b.float(), 1e-5).permute(0, 2, 1))
    err = (got.float() - want).abs().max().item()
    assert err < 2e-2 * max(1.0, want.abs().max().item() / 2.0), (n, hw, c1, c2, err)
print("GN_WS2_OK")
PY
echo "check rc=$?"; tail -3 gpurun_out/r2ay_check.log
VF_GN_WS=2 timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2ay_gn_ws2.txt 2>&1; echo "gn_ab ws=2 rc=$?"; cut -c1-41,102-170 gpurun_out/r2ay_gn_ws2.txt
VF_GN_WS=2 VF_GN_DEBUG_NOWAIT=1 timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2ay_gn_ws2_nowait.txt 2>&1; echo "nowait rc=$?"; cut -c1-41,102-170 gpurun_out/r2ay_gn_ws2_nowait.txt
