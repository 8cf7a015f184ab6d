#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/gputests_r2d.log
python bench.py > gpurun_out/r2_full.json 2> gpurun_out/r2_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_full.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), round(d["value"],1), {k:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items() if v["ms"]>0})
print("roofline", {k:d["roofline"][k] for k in ("kernel","achieved","frac","traffic")})
print("e2e", d.get("e2e"))
print("cpu_baseline", d.get("cpu_baseline"))
for k in ("query","xray","config1"):
    print(k, json.dumps(d.get(k))[:1200])
PY
tail -3 gpurun_out/r2_full.err
