#!/bin/bash
# round-2 GPU call AA: is the W-resident GEMM+GEGLU kernel bound by the depth of its A ring?  (library built with 2 stages)
mkdir -p gpurun_out
timeout 600 python benchmarks/bench_kernels.py --only gemm 2>&1 | grep "fused linear" | tee gpurun_out/r2aa_gemm_2stages.txt
