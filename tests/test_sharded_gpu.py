"""T11 on real GPUs: 2-rank NCCL sharded build == 1-GPU build (skipped on boxes with fewer than 2 GPUs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_build_matches_single_gpu():
    import point_cloud_viewer_b200 as pcv

    if pcv.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29741",
           os.path.join(ROOT, "scripts", "sharded_check.py"), "1e6", "20000", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
