#!/bin/bash
# round-2 GPU call Y: ncu of the round-robin attention kernel (with / without ordering) at 32 frame-branches
mkdir -p gpurun_out
for e in 1 2; do
  VF_ATTN_PP=$e timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp -c 1 -f -o gpurun_out/r2y_attn_pp$e python benchmarks/attn_once.py 32 1 > gpurun_out/r2y_ncu_$e.log 2>&1; echo "ncu $e rc=$?"
done
