set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass --launch-skip 2 --launch-count 1 -o gpurun_out/prof_r2_pass_v5 -f python bench.py --points 1e8 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
