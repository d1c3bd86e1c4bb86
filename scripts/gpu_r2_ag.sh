#!/bin/bash
# round-2 GPU call AG: single-tanh GELU epilogue (default) vs A&S erf (VF_GEMM_GELU=0), single-CTA and CTA-pair kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q -k "geglu or gemm" > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ag_tests.log
for g in 1 0; do
  echo "== VF_GEMM_GELU=$g"
  VF_GEMM_GELU=$g timeout 600 python benchmarks/bench_kernels.py --only gemm 2>&1 | grep "fused linear" | tee gpurun_out/r2ag_gemm_gelu$g.txt
  VF_GEMM_GELU=$g VF_GEMM_PAIR=1 timeout 120 python benchmarks/geglu_pair_check.py 2>&1 | tail -5 | tee gpurun_out/r2ag_pair_gelu$g.txt
done
