"""Phase timing of the attention softmax warps (needs a library built with VF_NVCC_EXTRA=-DVF_ATTN_TRACE):
    VF_NVCC_EXTRA=-DVF_ATTN_TRACE python -m vface_b200.build --force && python experiments/attn_trace.py [frames]"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops, _lib
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 96
lib = _lib.load()
lib.vf_attn_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda: torch.randn(frames, 4096, 320, device="cuda", generator=g).bfloat16()
q, k, v = mk(), mk(), mk()
out = torch.empty_like(q)
ops.attention(q, k, v, 8, out=out)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 48)()
lib.vf_attn_trace_read(buf, 1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); ops.attention(q, k, v, 8, out=out); ev1.record(); torch.cuda.synchronize()
lib.vf_attn_trace_read(buf, 0)
v_ = list(buf)
ctas, tiles = v_[8], v_[9]
names = ["prologue until S_0", "s_full waits", "S load + P hand-over", "row max", "exponentials", "p_empty wait", "P store + sums", "epilogue"]
print(f"kernel {ev0.elapsed_time(ev1):.3f} ms; traced CTAs {ctas}, tiles {tiles}")
tot = sum(v_[:8])
for n, c in zip(names, v_[:8]):
    per = c / tiles if "prologue" not in n and "epilogue" not in n else c / ctas
    unit = "clk/tile" if "prologue" not in n and "epilogue" not in n else "clk/CTA"
    print(f"  {n:24s} {100.0 * c / tot:5.1f} %   {per:9.1f} {unit}")
print(f"  total per CTA {tot / ctas:.0f} clk = {tot / tiles:.1f} clk per tile (one warp's timeline)")
for n, c in zip(["MMA thread: wait s_free", "MMA thread: issue QK (+k_full wait)", "MMA thread: wait v_full, p_full", "MMA thread: issue PV + commits"], v_[10:14]):
    print(f"  {n:38s} {c / tiles:9.1f} clk/tile")

print("per softmax warp (TMEM lane quarter), clk per tile:  s_full wait | S load | row max | exp | p_empty wait | store | total")
for w in range(4):
    r = v_[16 + 8 * w: 24 + 8 * w]
    print(f"  warp {w}: " + " | ".join(f"{r[i] / tiles:7.1f}" for i in (1, 2, 3, 4, 5, 6)) + f" | {sum(r[1:7]) / tiles:7.1f}")

ev = (ctypes.c_ulonglong * 128)()
lib.vf_attn_trace_read(ev, 2)
e = [list(ev[16 * t: 16 * t + 16]) for t in range(8)]
t0 = e[0][0]
print("event timeline of one CTA (cycles since s_full(20) passed), tiles 20..27:")
print("  tile | softmax: s_full passed, s_free arrived, exps done, tile done | issuer: s_free seen, QK(j+1) issued, p_full(j) seen, PV(j) issued | S(j+1) complete, PV(j) complete")
for t in range(8):
    f = lambda i: f"{e[t][i] - t0:7d}" if e[t][i] else "      -"
    print(f"  {20 + t:4d} | {f(0)} {f(1)} {f(2)} {f(3)} | {f(8)} {f(9)} {f(10)} {f(11)} | {f(12)} {f(13)}")
