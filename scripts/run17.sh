#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/gputests_r2c.log
b() { name=$1; shift; env "$@" python bench.py --roofline-only --steps 5 --warmup 3 > gpurun_out/r2_$name.json 2> gpurun_out/r2_$name.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["ms_per_step"],2), {k:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items() if v["ms"]>0})
PY
}
b q_default PCV_TIMING=1
b q_batch1 PCV_INGEST_BATCH=1
b q_generic PCV_PASS_GENERIC=1
grep "pcv timing" gpurun_out/r2_q_default.err | head -3
