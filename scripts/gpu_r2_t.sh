#!/bin/bash
# round-2 GPU call T: GEMM+GEGLU streaming kernel with / without the L2 prefetch of the next tile's A
mkdir -p gpurun_out
for e in 1 3; do
  echo "VF_GEMM_PREFETCH=$e"
  VF_GEMM_PREFETCH=$e timeout 900 python benchmarks/bench_kernels.py --only gemm --iters 40 2>&1 | grep "fused linear" | tee -a gpurun_out/r2t_gemm_pf$e.txt
done
