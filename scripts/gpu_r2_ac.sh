#!/bin/bash
# round-2 GPU call AC: resident GroupNorm with the wave-quantisation-aware slab plan -- tests + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_glue_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ac_tests.log
timeout 1200 python benchmarks/gn_ab.py > gpurun_out/r2ac_gn_ab.txt 2>&1; echo "gn_ab rc=$?"; cat gpurun_out/r2ac_gn_ab.txt
