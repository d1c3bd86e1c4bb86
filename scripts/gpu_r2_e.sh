#!/bin/bash
# round-2 GPU call E (2 GPUs): D1 on real ranks, the N=2 bench line (weak headline + 256-frame strong + shard check), new tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt 2>&1
echo "== new single-GPU tests"
timeout 900 python -m pytest tests/test_memory_guards_gpu.py tests/test_flow_producer_gpu.py -q > gpurun_out/r2e_newtests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2e_newtests.log
echo "== verify-shard on 2 ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --verify-shard > gpurun_out/r2e_verify_shard.json 2> gpurun_out/r2e_verify_shard.err; echo "rc=$?"
cat gpurun_out/r2e_verify_shard.json; tail -5 gpurun_out/r2e_verify_shard.err
echo "== multi-GPU pytest"
timeout 900 python -m pytest tests/test_frame_shard_gpu.py -q > gpurun_out/r2e_pytest_shard.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2e_pytest_shard.log
echo "== bench N=2"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2e_bench_n2.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","scaling","halo","clip256","shard_check")})
    print(d["e2e"], d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/r2e_bench_n2.err
