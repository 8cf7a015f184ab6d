"""Host orchestration of the GPU build (pass planning, leaf/split decisions, bucket tables, closed-form LOD
subsampling, output layout) run with the test-only CPU stand-ins for the kernels, against the oracle.  No GPU."""
import numpy as np
import pytest

import oracle_api as O
from parity import compare_trees
from tb_api import TbTree


def _data(rng, n, offset=(4.1e6, 6.6e5, 4.7e6), ncl=20):
    cen = rng.random((ncl, 3)) * [200, 200, 20]
    k = rng.integers(0, ncl, n)
    P = cen[k] + rng.normal(0, 1, (n, 3)) * rng.random((ncl, 1))[k] * 3
    P[: n // 10] = P[0]
    return P + offset


@pytest.mark.parametrize("n,maxpts,G,res", [(60000, 300, 3, 0.001), (60000, 300, 2, 0.001), (60000, 300, 1, 0.001), (20000, 40, 3, 1e-6),
                                             (80000, 700, 3, 1e-9), (5000, 100000, 3, 0.001), (1, 100000, 3, 0.001), (9, 1, 2, 0.5)])
def test_plan_matches_oracle(n, maxpts, G, res):
    rng = np.random.default_rng(n * 7 + G)
    P = _data(rng, n)
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    rgb = rng.integers(0, 255, n * 3, dtype=np.uint8)
    inten = rng.random(n).astype(np.float32)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, P.min(0), P.max(0), intensity=inten, max_points_per_node=maxpts)
    t = TbTree(x, y, z, rgb, res, P.min(0), P.max(0), maxpts, G, intensity=inten)
    compare_trees(ref, t)
    assert sum(v["num_points"] for v in t.nodes.values()) == n


def test_reference_scenario_and_G_independence():
    n = 100001
    x, y, z = np.zeros(n), np.zeros(n), np.zeros(n)
    x[-1], y[-1], z[-1] = -200.0, -40.0, 30.0
    rgb = np.tile(np.array([255, 0, 0], np.uint8), n)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), 1.0, (-200, -40, 0), (0, 0, 30))
    for G in (1, 2, 3):
        t = TbTree(x, y, z, rgb, 1.0, (-200, -40, 0), (0, 0, 30), 100000, G)
        assert {k: v["num_points"] for k, v in t.nodes.items()} == {"r": 12501, "r0": 0, "r4": 87500}
        compare_trees(ref, t)


def test_unsplittable_identical_points_reach_last_level():
    """generation.rs:137-147: > MAX points in a cell whose edge <= resolution stay together ("too small to be split")."""
    n = 5000
    rng = np.random.default_rng(2)
    P = rng.random((n, 3)) * 64.0
    P[:3000] = [10.123, 20.456, 30.789]
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    rgb = rng.integers(0, 255, n * 3, dtype=np.uint8)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), 0.25, (0, 0, 0), (64, 64, 64), max_points_per_node=1000)
    t = TbTree(x, y, z, rgb, 0.25, (0, 0, 0), (64, 64, 64), 1000, 3)
    compare_trees(ref, t)
    deepest = max(len(k) - 1 for k in t.nodes)
    assert deepest == 8  # edge 64 / 2^8 = 0.25 <= resolution: first unsplittable level
    assert max(v["num_points"] for k, v in t.nodes.items() if len(k) - 1 == deepest) > 1000


def test_aos_stride():
    rng = np.random.default_rng(4)
    P = np.ascontiguousarray(_data(rng, 20000))
    rgb = rng.integers(0, 255, 60000, dtype=np.uint8)
    flat = P.reshape(-1)
    t = TbTree(flat[0:], flat[1:], flat[2:], rgb, 0.01, P.min(0), P.max(0), 500, 3, stride=3)
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    compare_trees(O.build(x, y, z, rgb.reshape(-1, 3), 0.01, P.min(0), P.max(0), max_points_per_node=500), t)


def test_degenerate_numerators():
    rng = np.random.default_rng(17)
    n = 30000
    P = rng.random((n, 3))
    P[:3000] = 0.0
    P[3000:3500, 0] = -0.0
    P[3500:4000, 1] = 1e-200
    P[4000:4500, 2] = 5e-324
    P[4500:5000] = 1.0
    P[5000:5500] = 0.5
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    rgb = rng.integers(0, 255, n * 3, dtype=np.uint8)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), 1e-6, (0, 0, 0), (1, 1, 1), max_points_per_node=200)
    compare_trees(ref, TbTree(x, y, z, rgb, 1e-6, (0, 0, 0), (1, 1, 1), 200, 2))


def test_nan_inf_and_huge_coordinates():
    """Rust `as u8/u16` maps NaN to 0 and saturates; NaN survives Float32/Float64 codes (num::clamp passes it through).
    Host arithmetic (csrc/chain.h) against the oracle; the GPU twin is tests/test_build_gpu.py::test_wild_inputs_*."""
    rng = np.random.default_rng(23)
    n = 30000
    P = rng.random((n, 3)) * 100.0 + [4.1e6, 6.6e5, 4.7e6]
    bmin, bmax = P.min(0).copy(), P.max(0).copy()
    P[100] = [1e200, 6.6e5, 4.7e6]
    P[5000, 1] = -1e300
    P[9000, 2] = np.inf
    P[12000, 0] = -np.inf
    P[20000] = [np.nan, 6.6e5 + 1, 4.7e6 + 1]
    P[20001, 2] = 2.0 ** 399
    P[25000] = 0.0
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    rgb = rng.integers(0, 255, n * 3, dtype=np.uint8)
    for maxpts, res, G in ((300, 1e-4, 2), (2000, 1e-9, 3), (300, 1e-3, 1)):
        ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, max_points_per_node=maxpts)
        compare_trees(ref, TbTree(x, y, z, rgb, res, bmin, bmax, maxpts, G))
