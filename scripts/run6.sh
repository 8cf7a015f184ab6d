set -x
timeout 1200 python -m pytest tests/test_query_gpu.py tests/test_xray_full_size_gpu.py tests/test_full_size_gpu.py tests/test_cpp_host_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err
tail -5 gpurun_out/r2_b2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b2.json'))
print(d['ms_per_step'], d['wall_ms_per_step'], d['value'], d.get('library_event_ms_per_step'))
for k in ('frustum_query','xray','parity_check','cpu_baseline','config1','e2e'):
    print(k, json.dumps(d.get(k))[:1500])
PY
