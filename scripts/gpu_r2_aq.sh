#!/bin/bash
# round-2 GPU call AQ (2 GPUs): third-session state on real ranks: verify-shard, multi-GPU pytest, N=2 bench line, reference arm under torchrun
mkdir -p gpurun_out
echo "== verify-shard on 2 ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --verify-shard > gpurun_out/r2aq_verify_shard.json 2> gpurun_out/r2aq_verify_shard.err; echo "rc=$?"
grep verify_shard gpurun_out/r2aq_verify_shard.json; grep -i "error" gpurun_out/r2aq_verify_shard.err | head -5
echo "== multi-GPU pytest"
timeout 900 python -m pytest tests/test_frame_shard_gpu.py -q > gpurun_out/r2aq_pytest_shard.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2aq_pytest_shard.log
echo "== bench N=2"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2aq_bench_n2.json 2> gpurun_out/r2aq_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2aq_bench_n2.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","halo","shard_check","clip256")}); print(d["e2e"])
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2aq_bench_n2.err
