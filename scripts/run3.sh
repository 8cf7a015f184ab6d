set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
PCV_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_z2.json 2> gpurun_out/r2_z2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_z2.json'))
print(d['ms_per_step'], d['wall_ms_per_step'], d['value'])
for k,v in d['roofline']['kernels'].items(): print(k, v)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass --launch-skip 1 --launch-count 2 -o gpurun_out/prof_r2_pass_v2 -f python bench.py --points 1e8 --steps 1 --warmup 3 > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
