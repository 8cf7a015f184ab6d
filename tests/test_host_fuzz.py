"""Randomised host-logic check: csrc/build_host.hpp (pass planning, leaf / split decisions, closed-form subsampling, node
layout) driven through the test-only sequential backend must reproduce the oracle on arbitrary small clouds - duplicates,
flat and degenerate boxes, points on cube faces, every levels-per-pass setting."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle_api as O
from parity import compare_trees
from tb_api import TbTree


@st.composite
def clouds(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    n = draw(st.sampled_from([1, 2, 7, 8, 9, 63, 64, 65, 500, 3000, 20000]))
    kind = draw(st.sampled_from(["uniform", "clustered", "line", "plane", "duplicates", "lattice"]))
    scale = draw(st.sampled_from([1.0, 37.5, 1e4]))
    off = np.array(draw(st.sampled_from([(0.0, 0.0, 0.0), (4.1e6, 6.6e5, 4.7e6), (-3.0e5, 2.0e-3, 9.0e7)])))
    if kind == "uniform":
        P = rng.random((n, 3))
    elif kind == "clustered":
        c = rng.random((5, 3))
        P = np.clip(c[rng.integers(0, 5, n)] + rng.normal(0, 0.01, (n, 3)), 0, 1)
    elif kind == "line":
        P = np.outer(rng.random(n), [1.0, 0.5, 0.25])
    elif kind == "plane":
        P = rng.random((n, 3)) * [1, 1, 0]
    elif kind == "duplicates":
        P = rng.random((max(1, n // 50), 3))[rng.integers(0, max(1, n // 50), n)]
    else:
        P = rng.integers(0, 9, (n, 3)) / 8.0  # exactly on cell faces and centres
    P = P * scale + off
    rgb = rng.integers(0, 256, n * 3, dtype=np.uint8)
    inten = rng.random(n).astype(np.float32) if draw(st.booleans()) else None
    res = draw(st.sampled_from([scale * 0.5, scale * 1e-2, scale * 1e-5, scale * 1e-9]))
    maxpts = draw(st.sampled_from([1, 3, 40, 700, 100000]))
    G = draw(st.sampled_from([1, 2, 3]))
    pad = draw(st.sampled_from([0.0, 0.1]))  # the box is an argument: it may be larger than the data
    return P, rgb, inten, res, maxpts, G, pad


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(clouds())
def test_random_clouds_match_oracle(c):
    P, rgb, inten, res, maxpts, G, pad = c
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    ext = max(float((P.max(0) - P.min(0)).max()), 1e-6)
    bmin, bmax = P.min(0) - pad * ext, P.max(0) + pad * ext
    if not ((bmax - bmin).max() > 0):
        bmax = bmin + 1.0  # a single point: any positive box
    try:
        ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, intensity=inten, max_points_per_node=maxpts)
    except Exception:
        return  # configurations the oracle itself rejects (deeper than NodeId allows) are covered elsewhere
    t = TbTree(x, y, z, rgb, res, bmin, bmax, maxpts, G, intensity=inten)
    compare_trees(ref, t)
