#!/bin/bash
# round-2 GPU call S: GEMM+GEGLU W-resident kernel, L2 prefetch of the next tile's A on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "geglu or gemm" > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2s_tests.log
for e in 0 1; do
  echo "VF_GEMM_PREFETCH=$e"
  VF_GEMM_PREFETCH=$e timeout 900 python benchmarks/bench_kernels.py --only gemm 2>&1 | grep "fused linear" | tee -a gpurun_out/r2s_gemm_pf$e.txt
done
VF_GEMM_PREFETCH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_geglu -c 1 -f -o gpurun_out/r2s_geglu python benchmarks/kernel_once.py geglu_gemm bf16 1 > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc=$?"
