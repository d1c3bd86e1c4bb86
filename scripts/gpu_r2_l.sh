#!/bin/bash
# round-2 GPU call L: resident GroupNorm, trimmed sync path, 2 vs 3 CTAs/SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_glue_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2l_tests.log
VF_GN_RES_CTAS=3 timeout 900 python -m pytest tests/test_glue_kernels_gpu.py -x -q -k "group_norm or gn" > gpurun_out/r2l_tests3.log 2>&1; echo "tests(3 CTAs) rc=$?"; tail -3 gpurun_out/r2l_tests3.log
timeout 1200 python benchmarks/gn_ab.py > gpurun_out/r2l_gn_ab.txt 2>&1; echo "gn_ab rc=$?"; cat gpurun_out/r2l_gn_ab.txt
