#!/bin/bash
# round-2 GPU call AR: projection tests incl. the reference-golden case, smoke with the projection checks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_proj_gemm_gpu.py -x -q > gpurun_out/r2ar_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2ar_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ar_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2ar_smoke.log
