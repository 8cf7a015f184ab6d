"""BASELINE config 2 DATA against the oracle: the first 2e7 points of the SYNTH_GAUSS_CLUSTERS generator
bench.py builds at 1e9 — Gaussian clusters plus the eight blocks of 150 000 identical points (global indices 2^20 ..) that
run the "too small to be split" chain down to level 20 — built on the GPU and compared with the oracle bit for bit: node set,
counts, encodings, cubes, per-slot source index, colours and stored position codes.  tests/test_full_size_gpu.py covers the
full 1e9 through size-independent properties; this test is the direct comparison on the same distribution."""
import numpy as np
import pytest

import oracle_api as O
from parity import compare_trees

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("levels_per_pass", [2, 3])
def test_config2_sample_equals_oracle(levels_per_pass):
    import point_cloud_viewer_b200 as pcv

    n = 20_000_000
    kind = pcv.SYNTH_GAUSS_CLUSTERS
    x, y, z, rgb = pcv.synth_points_host(kind, 1, 0, n)  # bench.py's SEED
    bmin, bmax, res = pcv.synth_bbox(kind)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, num_threads=0)
    assert max(len(nm) - 1 for nm in ref.nodes) == 20  # the identical-point blocks reach the last level
    assert sum(1 for nm, m in ref.nodes.items() if len(nm) - 1 == 20 and m["num_points"] > 100000) == 8
    ctx = pcv.Context(0, levels_per_pass=levels_per_pass)
    tree = ctx.build_octree(x, y, z, rgb, res, bmin, bmax)
    compare_trees(ref, tree)
    # the device-side generator is the one bench.py uses at 1e9: same bits as the host one
    import torch

    xs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
    c = torch.empty(3 * n, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(kind, 1, 0, n, xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(), c.data_ptr())
    assert np.array_equal(xs[0].cpu().numpy(), x) and np.array_equal(xs[1].cpu().numpy(), y) and np.array_equal(xs[2].cpu().numpy(), z)
    assert np.array_equal(c.cpu().numpy(), rgb)
    tree.free()
    ctx.close()
