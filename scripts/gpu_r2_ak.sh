#!/bin/bash
# round-2 GPU call AK: projection kernel v2 (two epilogue groups, residual through identity MMAs, L2 prefetch): parity, A/B of the prefetch knob, ncu of QKV+LN
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_proj_gemm_gpu.py -x -q > gpurun_out/r2ak_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2ak_tests.log
for pf in 3 0; do
  VF_PROJ_PREFETCH=$pf timeout 300 python benchmarks/bench_proj.py > gpurun_out/r2ak_proj_pf$pf.txt 2>&1; echo "bench pf=$pf rc=$?"
  cat gpurun_out/r2ak_proj_pf$pf.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:proj3 -c 2 -f -o gpurun_out/r2ak_proj python benchmarks/proj_once.py > gpurun_out/r2ak_ncu.log 2>&1; echo "ncu rc=$?"
