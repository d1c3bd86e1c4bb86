#!/bin/bash
# round-2 GPU call I: fused GroupNorm -- parity tests, A/B table, where the small bf16 copies come from, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_glue_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2i_tests.log
timeout 1200 python benchmarks/gn_ab.py > gpurun_out/r2i_gn_ab.txt 2>&1; echo "gn_ab rc=$?"; cat gpurun_out/r2i_gn_ab.txt
timeout 600 python benchmarks/profile_step.py --top 12 --shapes aten::copy_,aten::_to_copy,aten::add,aten::contiguous,aten::clone > gpurun_out/r2i_profile_step.txt 2>&1; echo "profile rc=$?"; grep -E "^aten::" gpurun_out/r2i_profile_step.txt | head -40
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-clip256 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2i_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks")})
    for r in d["roofline_secondary"]["kernels"]:
        if r["kernel"].startswith("group_norm"): print(r["kernel"], round(r["ms_per_step"],3), round(r["frac"],3))
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2i_bench_n1.err
