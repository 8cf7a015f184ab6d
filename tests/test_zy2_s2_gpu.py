"""GPU parity of the S2-cell point cloud through the C ABI against the oracle's restatement of the s2 crate
arithmetic and of S2Splitter::write / S2Cells::nodes_in_location / CellUnion::contains (src/read_write/s2.rs,
src/s2_cells/mod.rs, src/geometry/s2_cell_union.rs).  Bit-exact: cell ids, cells, counts, per-cell order, survivors.
(The file name sorts last on purpose: these entry points were added after the last GPU session of their round.)"""
import numpy as np
import pytest

import s2_api as S

pytestmark = pytest.mark.gpu


def _slab(pcv, n):
    x, y, z, rgb = pcv.synth_points_host(pcv.SYNTH_SLAB_ECEF, 80293751232, 0, n)  # a 200 x 200 x 20 m slab at ECEF magnitude
    return x, y, z, rgb


def test_cell_ids(ctx):
    import point_cloud_viewer_b200 as pcv

    x, y, z, _ = _slab(pcv, 300_000)
    P = np.stack([x, y, z], 1)
    rng = np.random.default_rng(1)
    v = rng.normal(size=(50_000, 3))
    Q = np.concatenate([P, v / np.linalg.norm(v, axis=1)[:, None] * 6.37e6, v, np.array([[1, 1, 0], [0, 0, 0], [0.0, -0.0, 5.0], [1e308, 1e308, 1e308]], float)])
    for level in (30, 20, 9, 0):
        got = ctx.s2_cell_ids(*[np.ascontiguousarray(Q[:, k]) for k in range(3)], level)
        assert np.array_equal(got, S.oracle_cell_ids(Q, level)), level
    aos = np.ascontiguousarray(P[:1000])  # interleaved xyz
    got = ctx.s2_cell_ids(aos.ctypes.data, aos.ctypes.data + 8, aos.ctypes.data + 16, 20, stride=3, n=1000)
    assert np.array_equal(got, S.oracle_cell_ids(P[:1000], 20))


@pytest.mark.parametrize("level,n", [(20, 200_000), (23, 50_000), (14, 50_000)])
def test_split_equals_oracle(ctx, level, n):
    import point_cloud_viewer_b200 as pcv

    x, y, z, rgb = _slab(pcv, n)
    inten = ((np.arange(n) * 31) % 997).astype(np.float32)
    P = np.stack([x, y, z], 1)
    want = S.split(P, level)
    cloud = ctx.build_s2_cloud(x, y, z, rgb, inten, split_level=level)
    try:
        assert cloud.num_points == n and cloud.split_level == level and cloud.has_color and cloud.has_intensity
        assert np.array_equal(cloud.cell_ids, want["ids"]) and np.array_equal(cloud.cell_counts, want["counts"])
        assert np.array_equal(cloud.bbox_min, want["bmin"]) and np.array_equal(cloud.bbox_max, want["bmax"])
        o = 0
        rgb3 = rgb.reshape(-1, 3)
        for cid, cnt in zip(want["ids"], want["counts"]):
            idx = want["order"][o:o + int(cnt)]
            o += int(cnt)
            xyz, c, it, src = cloud.cell_data(cid)
            assert np.array_equal(src, idx), pcv.s2_token(cid)  # input order inside the cell
            assert np.array_equal(xyz, P[idx]) and np.array_equal(c, rgb3[idx]) and np.array_equal(it, inten[idx])
        # nodes_in_location + the filtered stream: AllPoints, the reference's own query shape (a level-20 cell and its
        # successor, point_cloud_test/src/queries.rs:49-53), a coarse and a mixed union
        allp = cloud.query_union(None)
        assert allp["total"] == n and np.array_equal(allp["src"], want["order"]) and np.array_equal(allp["xyz"], P[want["order"].astype(np.int64)])
        assert np.array_equal(cloud.cells_in_union(None), want["ids"])
        leaf = S.oracle_cell_ids(P, 30)
        centre = int(S.oracle_cell_ids(P[n // 2:n // 2 + 1], 20)[0])
        unions = [np.array([centre, S.orc().orc_s2_next(centre)], np.uint64), S.oracle_cell_ids(P[:3], 16), np.unique(np.concatenate([S.oracle_cell_ids(P[::5000], 22), S.oracle_cell_ids(P[7:8], 13)]))]
        for u in unions:
            un = S.normalize(u)
            cells_want = want["ids"][S.union_test(un, want["ids"])[1]]
            assert np.array_equal(cloud.cells_in_union(u), cells_want)
            inside = S.union_test(un, leaf)[0]
            keep = want["order"][inside[want["order"].astype(np.int64)]]
            got = cloud.query_union(u)
            assert got["total"] == len(keep) and np.array_equal(got["src"], keep), len(keep)
            assert np.array_equal(got["xyz"], P[keep.astype(np.int64)]) and np.array_equal(got["rgb"], rgb3[keep.astype(np.int64)])
            assert got["tested"] == int(want["counts"][np.isin(want["ids"], cells_want)].sum()) and len(keep) > 0
            assert np.array_equal(ctx.s2_union_contains(x, y, z, u), inside)
            part = cloud.query_union(u, cap=5)
            assert part["total"] == len(keep) and np.array_equal(part["src"], keep[:5])
        with pytest.raises(pcv._native.PcvError):
            cloud.cell_data(12345)
    finally:
        cloud.free()


def test_split_rejects_points_off_the_earth_and_handles_empty_input(ctx):
    import point_cloud_viewer_b200 as pcv

    x, y, z, rgb = _slab(pcv, 10_000)
    x2 = x.copy()
    x2[4321] *= 1.01
    with pytest.raises(pcv._native.PcvError) as e:
        ctx.build_s2_cloud(x2, y, z, rgb)
    assert "is not a valid ECEF point" in str(e.value)
    empty = ctx.build_s2_cloud(x[:0], y[:0], z[:0], rgb[:0])
    assert empty.num_points == 0 and empty.num_cells == 0 and empty.query_union(None)["total"] == 0
    empty.free()
    cloud = ctx.build_s2_cloud(x, y, z)  # positions only
    assert not cloud.has_color and cloud.query_union(None)["rgb"] is None and cloud.cell_counts.sum() == 10_000
    cloud.free()
