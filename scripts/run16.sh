#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/gputests_r2b.log
PCV_TIMING=1 python bench.py --roofline-only --steps 5 --warmup 3 > gpurun_out/r2_p.json 2> gpurun_out/r2_p.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_p.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], {k:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items()}, d.get("clocks"))
PY
grep "pcv timing" gpurun_out/r2_p.err | tail -2
