set -x
timeout 600 python scripts/records_selftest.py 300000 5000 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 scripts/sharded_check.py 1e6 20000 2 2>&1 | grep -E "OK|Error|error" | tail -8
PCV_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_m2.json 2> gpurun_out/r2_m2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_m2.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['parity_check'].get('equal'), d['full_size_check'], d['phases_ms'])
PY
grep -E "Error|error" gpurun_out/r2_m2.err | tail -5
