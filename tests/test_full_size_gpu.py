"""BASELINE.json full size (config 2: 1e9 Gaussian-cluster points, depth 20) through size-independent properties: point
conservation, the split rule, the source indices forming a permutation, colours travelling with their points, decoded
positions staying within the truncation bound, and frustum queries agreeing with a brute-force count."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Dev:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}


def test_config2_full_size_properties():
    import torch

    import point_cloud_viewer_b200 as pcv
    from point_cloud_viewer_b200 import _native as N

    free, _ = torch.cuda.mem_get_info()
    n = 1_000_000_000 if free > 150e9 else 200_000_000
    kind = pcv.SYNTH_GAUSS_CLUSTERS
    bmin, bmax, res = pcv.synth_bbox(kind)
    ctx = pcv.Context(0)
    x, y, z = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
    rgb = torch.empty(n * 3, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(kind, 1, 0, n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr())
    tree = ctx.build_octree(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), res, bmin, bmax, n=n, device=True)
    meta = tree.meta
    # (1) conservation (point_cloud_test/tests/main.rs:10-23) and the depth bound
    assert int(meta["num_points"].sum()) == n
    assert int(meta["level"].max()) == 20  # the 8 blocks of 150 000 identical points reach the last level
    # (2) split rule: more than MAX_POINTS_PER_NODE only where the cell cannot be split any further (or above such a cell)
    big = meta[meta["num_points"] > 100000]
    assert len(big) <= 8 * 21
    assert (big["level"] == 20).sum() == 8
    # (3) source indices are a permutation of 0..n-1: exact sum and sum of squares (mod 2^64)
    p_xyz, p_rgb, p_int, p_src = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    N.check(N.lib().pcv_octree_device_arrays(tree.h, C.byref(p_xyz), C.byref(p_rgb), C.byref(p_int), C.byref(p_src)))
    src = torch.as_tensor(_Dev(p_src.value, (n,), "<i4"), device="cuda")  # u32 viewed as i32 (n < 2^31)
    s64 = src.to(torch.int64)
    assert int(s64.sum()) == n * (n - 1) // 2
    want_sq = (n - 1) * n * (2 * n - 1) // 6
    assert int((s64 * s64).sum()) == ((want_sq + 2 ** 63) % 2 ** 64) - 2 ** 63  # int64 wrap-around arithmetic
    # (4) colours travel with their points
    out_rgb = torch.as_tensor(_Dev(p_rgb.value, (n, 3), "|u1"), device="cuda")
    step = 7
    sel = torch.arange(0, n, step, device="cuda")
    assert torch.equal(out_rgb[sel], rgb.view(n, 3)[s64[sel]])
    # (5) decoded positions of a sample of nodes stay within the truncation bound of their chain: every fix-point level a
    # point passed through truncates by less than its own code step, the coarsest fix-point level has a step below the
    # resolution (edge / resolution < 2^16 there) and the steps halve per level -> a few resolutions per axis in total;
    # f32 nodes are exact to 2^-23 of the edge
    rng = np.random.default_rng(0)
    pick = rng.choice(np.nonzero(meta["num_points"] > 0)[0], 60, replace=False)
    X = torch.stack([x, y, z], 1)
    for i in pick:
        m = meta[i]
        cnt, enc = int(m["num_points"]), int(m["enc"])
        bpc = {1: 1, 2: 2, 3: 4, 4: 8}[enc]
        raw = torch.as_tensor(_Dev(p_xyz.value + int(m["xyz_byte_offset"]), (cnt * 3 * bpc,), "|u1"), device="cuda")
        dt = {1: torch.uint8, 2: torch.int16, 3: torch.float32, 4: torch.float64}[enc]
        v = raw.view(dt).view(cnt, 3).to(torch.float64)
        if enc == 2:
            v = torch.where(v < 0, v + 65536.0, v)
        scale = {1: 255.0, 2: 65535.0, 3: 1.0, 4: 1.0}[enc]
        edge = float(m["cube"][3])
        dec = v / scale * edge + torch.tensor(m["cube"][:3].tolist(), device="cuda", dtype=torch.float64)
        po = int(m["point_offset"])
        err = (dec - X[s64[po:po + cnt]]).abs().max().item()
        tol = 4.0 * res + edge * 2.0 ** -20  # points stored in f32 nodes came up from fix-point nodes and carry their error
        assert err <= tol, (i, enc, err, tol)
    # (6) frustum queries: survivors of the octree query are inside the frustum by construction; compare the count with a
    # brute-force count over the ORIGINAL points (they may differ only by points within the quantisation error of a face)
    G = pcv.geometry
    locs = []
    for t in range(6):
        eye = bmin + rng.random(3) * (bmax - bmin)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        locs.append(G.frustum(G.Isometry(eye, q), G.Perspective.new_fov(1.0, 1.2, 0.1, 150.0)))
    counts, tested = tree.query_batch_device(locs)
    assert (tested >= counts).all() and counts.sum() > 0
    for loc, c in zip(locs, counts):
        M = torch.tensor(np.array(loc.clip_from_query).reshape(4, 4).T, device="cuda")  # column-major -> [r, c]
        acc = 0
        for s in range(0, n, 100_000_000):
            P = X[s:s + 100_000_000]
            h = P @ M[:3, :3].T + M[:3, 3]
            w = P @ M[3, :3] + M[3, 3]
            cl = h / w[:, None]
            acc += int(((cl.min(1).values > -1) & (cl.max(1).values < 1)).sum())
        assert abs(acc - int(c)) <= max(50, 0.002 * acc), (acc, int(c))
    tree.free()
    ctx.close()
