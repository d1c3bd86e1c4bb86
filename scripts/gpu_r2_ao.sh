#!/bin/bash
# round-2 GPU call AO: row-statistics hand-over (proj_in emits, QKV consumes): parity, A/B table, pipeline parity, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_proj_gemm_gpu.py -x -q > gpurun_out/r2ao_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2ao_tests.log
timeout 300 python benchmarks/bench_proj.py > gpurun_out/r2ao_proj.txt 2>&1; echo "bench rc=$?"
cat gpurun_out/r2ao_proj.txt
timeout 2400 python -m pytest tests/test_pipeline_gpu.py tests/test_hooks_gpu.py -x -q > gpurun_out/r2ao_tests2.log 2>&1; echo "tests2 rc=$?"; tail -4 gpurun_out/r2ao_tests2.log
timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-clip256 --no-elide-extra --no-graph-extra > gpurun_out/r2ao_bench.json 2> gpurun_out/r2ao_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2ao_bench.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks","gpu_launches")}); print(d["e2e"])
    for r in d["roofline_secondary"]["kernels"]:
        if "proj" in r["kernel"] or "layer_norm c=320" in r["kernel"]: print(r["kernel"], round(r["ms_per_step"],3), round(r["frac"],3))
except Exception as e: print("parse failed",e)
PY
