#!/bin/bash
# round-2 GPU call F: final-state validation on one GPU: suite, driver-style bench, reference arm, ncu launch list + attention capture
mkdir -p gpurun_out
echo "== suite"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2f_pytest.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2f_smoke.log
echo "== bench (driver-style)"
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2f_bench.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","clocks","gpu_launches")}); print(d["e2e"]); print(d["roofline"]); print(d["clip256"]); print(d["cpu_baseline"]); print(d["elide_dead_recon"])
for r in d["roofline_secondary"]["kernels"]: print(f"{r['kernel']:34s} {r['ms_per_step']:7.3f} ms {r['achieved']:8.1f} {r['unit']} {r['frac']:.3f}")
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "rc=$?"; cat gpurun_out/r2f_bench_ref.json | cut -c1-600
echo "== ncu launch list of the bench command"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-elide-extra --no-clip256 --no-secondary > gpurun_out/r2f_bench_short.json 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-elide-extra --no-clip256 --no-secondary > gpurun_out/r2f_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
echo "== ncu --set full, attention at the bench shape"
timeout 120 python benchmarks/attn_once.py 96 2 > gpurun_out/r2f_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -o gpurun_out/r2f_attn_tc python benchmarks/attn_once.py 96 2 > gpurun_out/r2f_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
