#!/bin/bash
# round-2 GPU call W: attention with early barrier probes (VF_ATTN_EARLY), correctness + timing A/B
mkdir -p gpurun_out
for e in 0 1 5 6 2 3; do
  echo "== VF_ATTN_EARLY=$e"
  VF_ATTN_EARLY=$e timeout 600 python benchmarks/attn_ab.py > gpurun_out/r2w_attn_early$e.txt 2>&1; echo "rc=$?"
  grep -E "timing|FAIL|\"ok\": false" gpurun_out/r2w_attn_early$e.txt | cut -c1-150
  tail -1 gpurun_out/r2w_attn_early$e.txt
done
