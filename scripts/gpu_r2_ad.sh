#!/bin/bash
# round-2 GPU call AD: default attention kernel with the exponentials as a bare MUFU stream (VF_ATTN_EARLY=7) vs default
mkdir -p gpurun_out
for e in 7 0 7 0; do
  echo "== VF_ATTN_EARLY=$e"
  VF_ATTN_EARLY=$e timeout 300 python benchmarks/attn_ab.py > gpurun_out/r2ad_attn_early$e.txt 2>&1; echo "rc=$?"
  grep -E "timing|FAIL|\"ok\": false" gpurun_out/r2ad_attn_early$e.txt | cut -c1-130
  tail -1 gpurun_out/r2ad_attn_early$e.txt
done
