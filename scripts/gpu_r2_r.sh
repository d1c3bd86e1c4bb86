#!/bin/bash
# round-2 GPU call R: GEMM+GEGLU W-resident kernel with four vs eight epilogue warps (staged TMA-store epilogue)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q -k "geglu or gemm" > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2r_tests.log
for e in 1 2; do
  echo "VF_GEMM_EPI_RES=$e"
  VF_GEMM_EPI_RES=$e timeout 900 python benchmarks/bench_kernels.py --only gemm 2>&1 | grep "fused linear" | tee -a gpurun_out/r2r_gemm_epi$e.txt
done
