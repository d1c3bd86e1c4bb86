#!/bin/bash
# round-2 GPU call G (2 GPUs): CUDA-graph step test + figure, verify-shard (fixed step count), multi-GPU pytest, N=2 line with the fp32 shard check
mkdir -p gpurun_out
echo "== graph test"
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -k cuda_graph > gpurun_out/r2g_graph_test.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2g_graph_test.log
echo "== verify-shard on 2 ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --verify-shard > gpurun_out/r2g_verify_shard.json 2> gpurun_out/r2g_verify_shard.err; echo "rc=$?"
grep verify_shard gpurun_out/r2g_verify_shard.json; grep -i "error" gpurun_out/r2g_verify_shard.err | head -5
echo "== multi-GPU pytest"
timeout 900 python -m pytest tests/test_frame_shard_gpu.py -q > gpurun_out/r2g_pytest_shard.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2g_pytest_shard.log
echo "== bench N=1 (graph figure)"
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-clip256 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2g_bench_n1.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","cuda_graph")})
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2g_bench_n1.err
echo "== bench N=2"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-clip256 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2g_bench_n2.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","halo","shard_check")})
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2g_bench_n2.err
