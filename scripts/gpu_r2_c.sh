#!/bin/bash
# round-2 GPU call C: stream kernel after the instruction diet (sum-based reference)
mkdir -p gpurun_out
run() { local name=$1; shift
  env "$@" timeout 200 python benchmarks/attn_ab.py --quick > gpurun_out/r2c_$name.log 2>&1
  echo "$name rc=$? fails=$(grep -c '"ok": false' gpurun_out/r2c_$name.log) $(grep timing gpurun_out/r2c_$name.log | head -1 | cut -c1-120)"
}
run old VF_ATTN_STREAM=0
run s3_l0 VF_ATTN_STREAM=1
run s3_l1 VF_ATTN_STREAM=1 VF_ATTN_LATE=1
run s3_l0_e1 VF_ATTN_STREAM=1 VF_ATTN_EMU=1
run s3_l0_e2 VF_ATTN_STREAM=1 VF_ATTN_EMU=2
run s3_l1_e1 VF_ATTN_STREAM=1 VF_ATTN_LATE=1 VF_ATTN_EMU=1
run s2_l0 VF_ATTN_STREAM=1 VF_ATTN_SBUF=2
echo "== d80 via stream (2 buffers)"
VF_ATTN_STREAM=2 timeout 200 python benchmarks/attn_ab.py > gpurun_out/r2c_stream2_full.log 2>&1; echo "rc=$?"; grep timing gpurun_out/r2c_stream2_full.log | cut -c1-120
echo "== kernel tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k attention > gpurun_out/r2c_pytest_attn.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2c_pytest_attn.log
