#!/bin/bash
# round-2 GPU call AW: gn_ws_kernel whole-group totals by the last arrival, two apply groups
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_glue_kernels_gpu.py -x -q -k "two_slab_pipeline and VF_GN_WS" > gpurun_out/r2aw_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2aw_tests.log
VF_GN_WS=1 timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2aw_gn_ws1.txt 2>&1; echo "gn_ab ws=1 rc=$?"; cut -c1-41,102-170 gpurun_out/r2aw_gn_ws1.txt
VF_GN_WS=1 VF_GN_DEBUG_NOWAIT=1 timeout 300 python benchmarks/gn_ab.py > gpurun_out/r2aw_gn_ws1_nowait.txt 2>&1; echo "nowait rc=$?"; cut -c1-41,102-170 gpurun_out/r2aw_gn_ws1_nowait.txt
