#!/bin/bash
# round-2 GPU call U: pipeline parity after the small-launch trims (cached bias sums, shared SiLU(emb), one-GEMM attn2 row) + bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_pipeline_gpu.py tests/test_hooks_gpu.py -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2u_tests.log
timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-clip256 > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2u_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks","gpu_launches")}); print(d["e2e"])
    for r in d["roofline_secondary"]["kernels"]:
        if "geglu" in r["kernel"] or "attn n=4096" in r["kernel"]: print(r["kernel"], round(r["ms_per_step"],3), round(r["frac"],3))
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2u_bench_n1.err
timeout 600 python experiments/bf16_error_by_block.py > gpurun_out/r2u_bf16_error.txt 2>&1; tail -8 gpurun_out/r2u_bf16_error.txt
