#!/bin/bash
# round-2 GPU call O: what does the inter-CTA wait of the resident GroupNorm cost? (no-wait timing experiment)
mkdir -p gpurun_out
timeout 1200 python benchmarks/gn_ab.py > gpurun_out/r2o_gn_ab.txt 2>&1; echo "gn_ab rc=$?"; cat gpurun_out/r2o_gn_ab.txt
