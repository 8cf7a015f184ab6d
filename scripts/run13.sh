#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/smi0.txt 2>&1
PCV_TIMING=2 PCV_RANK_KSTATS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 4 --warmup 3 --parity-points 2e6 > gpurun_out/r2_i.json 2> gpurun_out/r2_i.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_i.json
grep -E "pcv sharded|kstats" gpurun_out/r2_i.err | tail -8
# same with the ranks swapped onto the GPUs
CUDA_VISIBLE_DEVICES=1,0 PCV_TIMING=2 PCV_RANK_KSTATS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 4 --warmup 3 --parity-points 2e6 > gpurun_out/r2_j.json 2> gpurun_out/r2_j.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_j.json
grep -E "pcv sharded|kstats" gpurun_out/r2_j.err | tail -6
cat gpurun_out/topo.txt | head -12
