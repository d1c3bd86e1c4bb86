#!/bin/bash
# round-2 GPU call X: round-robin attention arrangement (vf_attn_pp.cu): correctness + timing, with / without the ordering
mkdir -p gpurun_out
for e in 1 2 3 0; do
  echo "== VF_ATTN_PP=$e"
  VF_ATTN_PP=$e timeout 180 python benchmarks/attn_ab.py --quick > gpurun_out/r2x_attn_pp$e.txt 2>&1; echo "rc=$?"
  grep -E "timing|FAIL|\"ok\": false" gpurun_out/r2x_attn_pp$e.txt | cut -c1-150
  tail -2 gpurun_out/r2x_attn_pp$e.txt | cut -c1-300
done
nvidia-smi --query-gpu=name,memory.used --format=csv
