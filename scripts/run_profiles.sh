set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputests_r2_final.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --points 1e8 --steps 2 --warmup 3 --frusta 100 --cpu-points 1e6 --ply-points 1e7 --xray-px 1024 > gpurun_out/launches_r2_bench.json 2> gpurun_out/launches_r2.err
cap() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 --launch-count $4 -o gpurun_out/tmp_$1 -f python scripts/profile_driver.py 1e8 > gpurun_out/tmp_$1.log 2>&1
  ncu -i gpurun_out/tmp_$1.ncu-rep --page details > gpurun_out/ncu_r2_$1_details.txt 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/tmp_$1.ncu-rep > gpurun_out/ncu_r2_$1.txt 2>&1
  rm -f gpurun_out/tmp_$1.ncu-rep gpurun_out/tmp_$1.log
}
cap k_pass "k_pass" 0 3
cap k_ingest "k_ingest" 0 1
cap k_place "k_place" 0 1
cap k_cull_fused "k_cull_fused" 0 1
cap k_xray "k_xray_bin|k_xray_subtile" 0 3
cap k_sub "k_sub_level|k_sub_layout" 20 2
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python scripts/records_selftest.py 100000 5000 > gpurun_out/sanitizer_racecheck_records.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_records.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_sharded_native_gpu.py -m gpu -x -q -k "SLAB or 1e-06" > gpurun_out/sanitizer_memcheck_sharded_native.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_sharded_native.log
du -sh gpurun_out
