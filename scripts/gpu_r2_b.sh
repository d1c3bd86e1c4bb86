#!/bin/bash
# round-2 GPU call B: stream-kernel variants (S buffers x late arrive x emu), correctness of the default, one ncu capture
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 200 python benchmarks/attn_ab.py --quick > gpurun_out/r2b_$name.log 2>&1
  echo "$name rc=$? $(grep timing gpurun_out/r2b_$name.log | head -1)"
}
run old VF_ATTN_STREAM=0
run s2_l0 VF_ATTN_STREAM=1 VF_ATTN_SBUF=2 VF_ATTN_LATE=0
run s2_l1 VF_ATTN_STREAM=1 VF_ATTN_SBUF=2 VF_ATTN_LATE=1
run s3_l0 VF_ATTN_STREAM=1 VF_ATTN_SBUF=3 VF_ATTN_LATE=0
run s3_l1 VF_ATTN_STREAM=1 VF_ATTN_SBUF=3 VF_ATTN_LATE=1
run s3_l1_e1 VF_ATTN_STREAM=1 VF_ATTN_SBUF=3 VF_ATTN_LATE=1 VF_ATTN_EMU=1
run s3_l1_e2 VF_ATTN_STREAM=1 VF_ATTN_SBUF=3 VF_ATTN_LATE=1 VF_ATTN_EMU=2
echo "== full A/B of the default"
timeout 300 python benchmarks/attn_ab.py > gpurun_out/r2b_default_full.log 2>&1; echo "rc=$?"
grep -c '"ok": true' gpurun_out/r2b_default_full.log; grep -c '"ok": false' gpurun_out/r2b_default_full.log; grep timing gpurun_out/r2b_default_full.log
echo "== ncu (default kernel, 16 frames)"
timeout 120 python benchmarks/attn_once.py 16 2 > gpurun_out/r2b_once_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_stream -s 1 -c 1 -o gpurun_out/r2b_attn_stream python benchmarks/attn_once.py 16 2 > gpurun_out/r2b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu.log
