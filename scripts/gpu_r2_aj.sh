#!/bin/bash
# round-2 GPU call AJ: first run of the tcgen05 projection kernel (parity, then the A/B against the library path)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_proj_gemm_gpu.py -x -q > gpurun_out/r2aj_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2aj_tests.log
timeout 300 python benchmarks/bench_proj.py > gpurun_out/r2aj_proj.txt 2>&1; echo "bench rc=$?"
cat gpurun_out/r2aj_proj.txt
