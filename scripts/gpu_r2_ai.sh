#!/bin/bash
# round-2 GPU call AI: ncu of the W-resident GEMM+GEGLU kernel with the single-tanh epilogue
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_geglu -c 1 -f -o gpurun_out/r2ai_geglu python benchmarks/kernel_once.py geglu_gemm bf16 1 > gpurun_out/r2ai_ncu.log 2>&1; echo "ncu rc=$?"
