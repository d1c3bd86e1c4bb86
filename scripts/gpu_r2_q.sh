#!/bin/bash
# round-2 GPU call Q: GEMM+GEGLU epilogue through a swizzled staging buffer + TMA store
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q -k "geglu or gemm" > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2q_tests.log
timeout 900 python benchmarks/bench_kernels.py --only gemm > gpurun_out/r2q_gemm.txt 2>&1; echo "bench rc=$?"; cat gpurun_out/r2q_gemm.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_geglu -c 1 -f -o gpurun_out/r2q_geglu python benchmarks/kernel_once.py geglu_gemm bf16 1 > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
