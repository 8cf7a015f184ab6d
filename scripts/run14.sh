#!/bin/bash
mkdir -p gpurun_out
PCV_TRACE_ALLOC=1 PCV_TIMING=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 3 --warmup 3 --parity-points 2e6 > gpurun_out/r2_k.json 2> gpurun_out/r2_k.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_k.json
grep -E "pcv sharded|pcv ingest|pcv dmalloc" gpurun_out/r2_k.err | tail -24
