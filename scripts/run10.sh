#!/bin/bash
# 2-GPU box: native (C-orchestrated) sharded build tests + fused exchange pass + bench with phase timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_native_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/native_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 scripts/sharded_check.py 1e6 20000 2 2>&1 | grep -E "OK|Error|error|assert" | tail -12 | tee gpurun_out/sharded_check2.log
PCV_TIMING=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_f2.json 2> gpurun_out/r2_f2.err
tail -c 1500 gpurun_out/r2_f2.json
grep -E "pcv sharded|pcv timing" gpurun_out/r2_f2.err | tail -12
