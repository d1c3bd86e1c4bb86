#!/bin/bash
# round-2 GPU call H1: CUDA-graph step test + figure
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -k cuda_graph > gpurun_out/r2h_graph_test.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2h_graph_test.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-clip256 --no-secondary > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2h_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","cuda_graph","elide_dead_recon")})
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2h_bench_n1.err
