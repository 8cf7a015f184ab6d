set -x
timeout 900 python -m pytest tests/test_build_gpu.py tests/test_config2_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --roofline-only > gpurun_out/r2_p1.json 2> gpurun_out/r2_p1.err
PCV_NO_POW2=1 timeout 600 python bench.py --steps 5 --warmup 3 --roofline-only > gpurun_out/r2_p1b.json 2>> gpurun_out/r2_p1.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_p1.json','gpurun_out/r2_p1b.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['value'], d.get('library_event_ms_per_step'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items() if v['ms']>0})
PY
