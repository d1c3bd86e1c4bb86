#!/bin/bash
# round-2 GPU call N: resident GroupNorm, 320 vs 512 threads per CTA
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_glue_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2n_tests.log
timeout 1200 python benchmarks/gn_ab.py > gpurun_out/r2n_gn_ab.txt 2>&1; echo "gn_ab rc=$?"; cat gpurun_out/r2n_gn_ab.txt
VF_GN_FUSED=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ -c 1 -f -o gpurun_out/r2n_gn_resident python benchmarks/kernel_once.py gn bf16 1 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
