#!/bin/bash
# round-2 GPU call AF: CTA-pair GEMM+GEGLU (tcgen05.mma.cta_group::2) -- correctness and time vs the single-CTA kernel
mkdir -p gpurun_out
for e in 1 0; do
  VF_GEMM_PAIR=$e timeout 120 python benchmarks/geglu_pair_check.py > gpurun_out/r2af_pair$e.txt 2>&1; echo "pair=$e rc=$?"; tail -6 gpurun_out/r2af_pair$e.txt | cut -c1-200
done
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
