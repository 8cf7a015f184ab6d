#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 4 --warmup 3 --parity-points 2e6 > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_$name.json
  grep -E "pcv sharded" gpurun_out/r2_$name.err | tail -8
}
run g_nosampler PCV_TIMING=2 PCV_NO_SAMPLER=1
run g_sampler PCV_TIMING=2
run g_pyorch PCV_TIMING=1 PCV_PY_ORCH=1
