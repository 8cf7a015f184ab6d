#!/bin/bash
# usage: multi_gpu_bench.sh N  (EXTRA_ENV="PCV_FUSED_PASS=1" ... for variants)  - the sharded bench on N GPUs with per-rank phase timings
N=${1:-8}
mkdir -p gpurun_out
env PCV_TIMING=2 ${EXTRA_ENV} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_n$N.json 2> gpurun_out/r2_n$N.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_n$N.json
grep -o '"parity_check": {[^}]*}' gpurun_out/r2_n$N.json | cut -c1-200
grep -o '"full_size_check": {[^}]*}' gpurun_out/r2_n$N.json
grep -E "pcv sharded" gpurun_out/r2_n$N.err | tail -$N
tail -3 gpurun_out/r2_n$N.err | cut -c1-300
