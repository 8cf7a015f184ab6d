set -x
nproc; free -g | head -2; df -h /dev/shm | tail -1; nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
PCV_TIMING=1 python bench.py --steps 3 --warmup 3 --levels-per-pass 2 > gpurun_out/r2_base_g2.json 2> gpurun_out/r2_base_g2.err
PCV_TIMING=1 python bench.py --steps 3 --warmup 3 --levels-per-pass 3 > gpurun_out/r2_base_g3.json 2> gpurun_out/r2_base_g3.err
tail -c 600 gpurun_out/r2_base_g3.json
