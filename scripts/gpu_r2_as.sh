#!/bin/bash
# round-2 GPU call AS: GroupNorm with two slabs per CTA in flight (gn_resident2_kernel): parity (every process under its own timeout: a
# sample barrier that never completes must not hang the box), then the A/B against the one-slab kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_glue_kernels_gpu.py tests/test_memory_guards_gpu.py -x -q -k "group_norm or gn or persistent or memory or guard" > gpurun_out/r2as_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2as_tests.log
for pipe in 1 0; do
  VF_GN_PIPE=$pipe timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2as_gn_pipe$pipe.txt 2>&1; echo "gn_ab pipe=$pipe rc=$?"; cat gpurun_out/r2as_gn_pipe$pipe.txt
done
VF_GN_PIPE=1 VF_GN_DEBUG_NOWAIT=1 timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2as_gn_pipe1_nowait.txt 2>&1; echo "nowait rc=$?"; cat gpurun_out/r2as_gn_pipe1_nowait.txt
