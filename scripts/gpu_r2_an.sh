#!/bin/bash
# round-2 GPU call AN: projection kernel in the model: full GPU suite, bench A/B (VF_PROJ_TC=0 vs 1), bf16 error by block
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/r2an_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2an_tests.log
for tc in 1 0; do
VF_PROJ_TC=$tc timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-clip256 --no-elide-extra --no-graph-extra > gpurun_out/r2an_bench_tc$tc.json 2> gpurun_out/r2an_bench_tc$tc.err; echo "bench tc=$tc rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2an_bench_tc$tc.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks","gpu_launches")}); print(d["e2e"])
    for r in d["roofline_secondary"]["kernels"]:
        if "proj" in r["kernel"] or "layer_norm c=320" in r["kernel"]: print(r["kernel"], round(r["ms_per_step"],3), round(r["frac"],3))
except Exception as e: print("parse failed",e)
PY
tail -2 gpurun_out/r2an_bench_tc$tc.err
done
timeout 600 python experiments/bf16_error_by_block.py > gpurun_out/r2an_bf16_error.txt 2>&1; tail -5 gpurun_out/r2an_bf16_error.txt
