#!/bin/bash
# round-2 GPU call AU: warp-specialised GroupNorm (gn_ws_kernel, VF_GN_WS=1): parity under a timeout, then the A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_glue_kernels_gpu.py -x -q -k "two_slab_pipeline" > gpurun_out/r2au_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2au_tests.log
VF_GN_WS=1 timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2au_gn_ws1.txt 2>&1; echo "gn_ab ws=1 rc=$?"; cat gpurun_out/r2au_gn_ws1.txt
VF_GN_WS=1 VF_GN_DEBUG_NOWAIT=1 timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2au_gn_ws1_nowait.txt 2>&1; echo "nowait rc=$?"; cat gpurun_out/r2au_gn_ws1_nowait.txt
