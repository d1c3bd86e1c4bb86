#!/bin/bash
# round-2 GPU call A: new attention kernel A/B, then the full GPU test suite, then a short bench
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
echo "== attn A/B: old arrangement" 
VF_ATTN_STREAM=0 timeout 300 python benchmarks/attn_ab.py > gpurun_out/r2a_attn_old.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/r2a_attn_old.log
echo "== attn A/B: stream"
VF_ATTN_STREAM=1 timeout 300 python benchmarks/attn_ab.py > gpurun_out/r2a_attn_stream.log 2>&1; RC=$?; echo "rc=$RC"
tail -30 gpurun_out/r2a_attn_stream.log
if [ $RC -ne 0 ]; then
  echo "stream kernel failed: falling back to VF_ATTN_STREAM=0 for the rest"
  export VF_ATTN_STREAM=0
else
  for emu in 1 2; do
    VF_ATTN_STREAM=1 VF_ATTN_EMU=$emu timeout 200 python benchmarks/attn_ab.py --quick > gpurun_out/r2a_attn_stream_emu$emu.log 2>&1; echo "emu$emu rc=$?"
    grep timing gpurun_out/r2a_attn_stream_emu$emu.log
  done
fi
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "rc=$?"
tail -15 gpurun_out/r2a_pytest.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "rc=$?"
cat gpurun_out/r2a_bench.json | head -c 6000
tail -5 gpurun_out/r2a_bench.err
