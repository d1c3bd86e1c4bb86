#!/bin/bash
# round-2 GPU call AZ: final tree after the GroupNorm experiments (default path unchanged): full GPU suite, smoke, driver-style bench
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/r2az_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2az_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2az_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2az_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2az_bench_n1.json 2> gpurun_out/r2az_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2az_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks","gpu_launches")}); print(d["e2e"]); print({k:d["roofline"][k] for k in ("achieved","frac","frac_of_mufu_ex2_ceiling","share_of_step")})
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2az_bench_n1.err
