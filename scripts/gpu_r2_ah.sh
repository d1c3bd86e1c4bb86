#!/bin/bash
# round-2 GPU call AH: pipeline parity with the single-tanh GELU epilogue + per-block bf16 error + short bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_pipeline_gpu.py tests/test_hooks_gpu.py tests/test_vae_decoder_gpu.py -x -q > gpurun_out/r2ah_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ah_tests.log
timeout 600 python experiments/full_size_errors.py > gpurun_out/r2ah_full_size_errors.txt 2>&1; echo "errors rc=$?"; tail -12 gpurun_out/r2ah_full_size_errors.txt
VF_GEMM_GELU=0 timeout 600 python experiments/full_size_errors.py > gpurun_out/r2ah_full_size_errors_as.txt 2>&1; echo "errors(A&S) rc=$?"; tail -12 gpurun_out/r2ah_full_size_errors_as.txt
timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-clip256 --no-graph-extra --no-elide-extra > gpurun_out/r2ah_bench_n1.json 2> gpurun_out/r2ah_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2ah_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks")}); print(d["e2e"]["value"])
    for r in d["roofline_secondary"]["kernels"]:
        if "geglu" in r["kernel"]: print(r["kernel"], round(r["ms_per_step"],3), round(r["frac"],3))
except Exception as e: print("parse failed",e)
PY
