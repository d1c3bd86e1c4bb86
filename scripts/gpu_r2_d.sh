#!/bin/bash
# round-2 GPU call D: wide-head attention + full first-stage decoder, whole GPU suite with the final defaults, then memcheck
mkdir -p gpurun_out
echo "== new tests first"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_vae_decoder_gpu.py -x -q -k "wide_head or streamed or full_ddconfig or psnr_full" > gpurun_out/r2d_new.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2d_new.log
echo "== whole suite"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2d_pytest.log
echo "== sanitize driver, plain"
timeout 300 python benchmarks/sanitize_kernels.py > gpurun_out/r2d_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python benchmarks/sanitize_kernels.py > gpurun_out/r2d_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -12 gpurun_out/r2d_memcheck.log
