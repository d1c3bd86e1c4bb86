#!/bin/bash
# round-2 GPU call AB (2 GPUs): sharded == unsharded on real ranks, multi-GPU pytest, N=2 bench line (weak + 256-frame strong),
# reference arm at N=2; plus the new GroupNorm stress test
mkdir -p gpurun_out
echo "== glue tests (incl. the persistent-grid GroupNorm stress test)"
timeout 900 python -m pytest tests/test_glue_kernels_gpu.py -q > gpurun_out/r2ab_glue.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2ab_glue.log
echo "== verify-shard on 2 ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --verify-shard > gpurun_out/r2ab_verify_shard.json 2> gpurun_out/r2ab_verify_shard.err; echo "rc=$?"
grep verify_shard gpurun_out/r2ab_verify_shard.json | cut -c1-600; grep -i "error" gpurun_out/r2ab_verify_shard.err | head -5
echo "== multi-GPU pytest"
timeout 900 python -m pytest tests/test_frame_shard_gpu.py -q > gpurun_out/r2ab_pytest_shard.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2ab_pytest_shard.log
echo "== bench N=2 (driver style)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2ab_bench_n2.json 2> gpurun_out/r2ab_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2ab_bench_n2.json") if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","halo","shard_check","clip256")})
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2ab_bench_n2.err
