set -x
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 scripts/sharded_check.py 5e5 20000 2 2>&1 | grep -E "OK|Error|error" | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_m$N.json 2> gpurun_out/r2_m$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_m$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['parity_check'].get('equal'), d['full_size_check'], d['phases_ms'])
PY
grep -E "Error|error" gpurun_out/r2_m$N.err | tail -5
