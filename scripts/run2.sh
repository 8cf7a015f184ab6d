set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
PCV_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_z1.json 2> gpurun_out/r2_z1.err
tail -c 1500 gpurun_out/r2_z1.json; grep "pcv timing" gpurun_out/r2_z1.err | tail -3
