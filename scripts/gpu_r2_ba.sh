#!/bin/bash
# round-2 GPU call BA (4 GPUs): last tree under torchrun on 4 ranks: weak-scaling line with shard check and the 256-frame clip
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/r2ba_bench_n4.json 2> gpurun_out/r2ba_bench_n4.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2ba_bench_n4.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","halo","shard_check","clip256")}); print(d["e2e"])
except Exception as e: print("parse failed",e)
PY
tail -3 gpurun_out/r2ba_bench_n4.err
