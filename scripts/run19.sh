#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2_final_1e9.json 2> gpurun_out/bench_r2_final_1e9.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_final_1e9.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2), round(d["value"],1), {k:round(v["ms"],2) for k,v in d["roofline"]["kernels"].items() if v["ms"]>0}, d["clocks"])
print("roofline", {k:d["roofline"][k] for k in ("kernel","achieved","frac","traffic")}, "whole", d["roofline"]["whole_build"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
fq=d["frustum_query"]
for k in fq: print(k, fq[k]["ms_device"], fq[k]["Mpoints_per_s_tested"], fq[k]["roofline"]["frac"])
print("xray", d["xray"]["ms_device"], d["xray"]["Mpoints_per_s"], d["xray"]["roofline"]["frac"])
print("parity", d.get("parity_check"))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_place --launch-skip 0 --launch-count 1 -o gpurun_out/tmp_k_place -f python scripts/profile_driver.py 1e8 > gpurun_out/tmp_k_place.log 2>&1
ncu -i gpurun_out/tmp_k_place.ncu-rep --page details > gpurun_out/ncu_r2_k_place_details.txt 2>/dev/null
python scripts/ncu_summary.py gpurun_out/tmp_k_place.ncu-rep > gpurun_out/ncu_r2_k_place.txt 2>&1
rm -f gpurun_out/tmp_k_place.ncu-rep gpurun_out/tmp_k_place.log
head -9 gpurun_out/ncu_r2_k_place.txt
