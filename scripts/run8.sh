set -x
PCV_TIMING=1 timeout 600 python scripts/dbg_shard.py 2>&1 | grep -v "pcv timing" | tail -8
timeout 600 python scripts/records_selftest.py 300000 5000 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/records_selftest.py 100000 5000 2>&1 | grep -vE "^=+$" | tail -12
