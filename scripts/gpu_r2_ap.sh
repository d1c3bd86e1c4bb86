#!/bin/bash
# round-2 GPU call AP: final-state validation of the third session: full GPU suite, smoke, driver-style bench, reference arm,
# ncu launch list of the bench command, ncu --set full of the four forms of the projection kernel
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/r2ap_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ap_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ap_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2ap_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ap_bench_n1.json 2> gpurun_out/r2ap_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2ap_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks","gpu_launches")}); print(d["e2e"]); print({k:d["roofline"][k] for k in ("achieved","frac","frac_of_mufu_ex2_ceiling","share_of_step")})
    print(d.get("elide_dead_recon")); print(d.get("clip256")); print(d.get("cuda_graph"))
    for r in d["roofline_secondary"]["kernels"]:
        print(" ", r["kernel"], round(r["ms_per_step"],3), round(r["frac"],3), round(r.get("frac_of_floor",0),3))
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2ap_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2ap_bench_ref.json 2> gpurun_out/r2ap_bench_ref.err; echo "ref arm rc=$?"; tail -c 600 gpurun_out/r2ap_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2ap_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-clip256 --no-secondary --no-e2e --no-elide-extra --no-graph-extra > gpurun_out/r2ap_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:proj3 -c 4 -f -o gpurun_out/r2ap_proj python benchmarks/proj_once.py > gpurun_out/r2ap_ncu.log 2>&1; echo "ncu rc=$?"
