#!/bin/bash
# round-2 GPU call J: resident GroupNorm -- parity tests, A/B table, ncu of the three forms, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_glue_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2j_tests.log
timeout 1200 python benchmarks/gn_ab.py > gpurun_out/r2j_gn_ab.txt 2>&1; echo "gn_ab rc=$?"; cat gpurun_out/r2j_gn_ab.txt
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-clip256 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2j_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks")})
    for r in d["roofline_secondary"]["kernels"]:
        if r["kernel"].startswith("group_norm"): print(r["kernel"], round(r["ms_per_step"],3), round(r["frac"],3))
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2j_bench_n1.err
for f in 2 1; do
  VF_GN_FUSED=$f timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ -c 2 -f -o gpurun_out/r2j_gn_fused$f python benchmarks/kernel_once.py gn bf16 2 > gpurun_out/r2j_ncu_$f.log 2>&1; echo "ncu $f rc=$?"
done
