#!/bin/bash
# round-2 GPU call AE: new full-size property tests + experimental attention knobs test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_glue_kernels_gpu.py tests/test_kernels_gpu.py -x -q -k "properties or experimental or counter_reuse" > gpurun_out/r2ae_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2ae_tests.log
