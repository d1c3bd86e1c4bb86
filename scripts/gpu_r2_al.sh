#!/bin/bash
# round-2 GPU call AL: projection kernel v3 (128-byte staging rows): parity + A/B table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_proj_gemm_gpu.py -x -q > gpurun_out/r2al_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2al_tests.log
timeout 300 python benchmarks/bench_proj.py > gpurun_out/r2al_proj.txt 2>&1; echo "bench rc=$?"
cat gpurun_out/r2al_proj.txt
