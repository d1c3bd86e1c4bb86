#!/bin/bash
# round-2 GPU call V: host-side overhead of sample() outside the steps
mkdir -p gpurun_out
timeout 900 python experiments/e2e_overhead.py > gpurun_out/r2v_e2e_overhead.txt 2>&1; echo "rc=$?"; head -60 gpurun_out/r2v_e2e_overhead.txt | cut -c1-160
