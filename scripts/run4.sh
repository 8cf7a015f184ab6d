set -x
timeout 900 python -m pytest tests/test_build_gpu.py tests/test_config2_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_z5.json 2> gpurun_out/r2_z5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_z5.json'))
print(d['ms_per_step'], d['wall_ms_per_step'], d['value'])
for k,v in d['roofline']['kernels'].items(): print(k, v)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches_r2_z5.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
