set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 scripts/sharded_check.py 1e6 20000 2 2>&1 | tail -8
PCV_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_m2.json 2> gpurun_out/r2_m2.err
tail -c 2500 gpurun_out/r2_m2.json; grep -E "pcv sharded|Error|error" gpurun_out/r2_m2.err | tail -5
