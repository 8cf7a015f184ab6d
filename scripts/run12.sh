#!/bin/bash
mkdir -p gpurun_out
python bench.py --roofline-only --steps 3 --warmup 3 > gpurun_out/r2_h0.json 2> gpurun_out/r2_h0.err
PCV_FIRST_INDEX=1e9 python bench.py --roofline-only --steps 3 --warmup 3 > gpurun_out/r2_h1.json 2> gpurun_out/r2_h1.err
python - <<'PY'
import json
for f in ("h0","h1"):
    d=json.loads(open("gpurun_out/r2_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], {k:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, d.get("clocks"))
PY
