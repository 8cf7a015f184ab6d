"""The on-disk form of an S2-cell cloud (read_write/s2.rs:127-145, raw.rs, s2_cells/mod.rs:77-147): per-cell files read back with
numpy, meta.pb parsed by python-protobuf, and the load -> query round trip.  (Sorts last: added after the round's last GPU session.)"""
import os

import numpy as np
import pytest

import s2_api as S

pytestmark = pytest.mark.gpu


def test_write_dir_and_load_dir(ctx, tmp_path):
    import point_cloud_viewer_b200 as pcv
    from proto_meta import Meta

    n = 60_000
    x, y, z, rgb = pcv.synth_points_host(pcv.SYNTH_SLAB_ECEF, 80293751232, 0, n)
    inten = ((np.arange(n) * 13) % 511).astype(np.float32)
    P = np.stack([x, y, z], 1)
    rgb3 = rgb.reshape(-1, 3)
    want = S.split(P, 21)
    cloud = ctx.build_s2_cloud(x, y, z, rgb, inten, split_level=21)
    d = str(tmp_path / "s2")
    cloud.write_dir(d)
    m = Meta.FromString(open(os.path.join(d, "meta.pb"), "rb").read())
    assert m.version == 13 and m.WhichOneof("data") == "s2"
    assert {c.id: c.num_points for c in m.s2.cells} == {int(i): int(c) for i, c in zip(want["ids"], want["counts"])}
    assert {a.name: a.data_type for a in m.s2.attributes} == {"color": 27, "intensity": 11}
    bb = m.bounding_box
    assert [bb.min.x, bb.min.y, bb.min.z] == list(want["bmin"]) and [bb.max.x, bb.max.y, bb.max.z] == list(want["bmax"])
    files = set(os.listdir(d))
    assert files == {"meta.pb"} | {pcv.s2_token(i) + e for i in want["ids"] for e in (".xyz", ".rgb", ".intensity")}
    o = 0
    for cid, cnt in zip(want["ids"], want["counts"]):
        idx = want["order"][o:o + int(cnt)].astype(np.int64)
        o += int(cnt)
        stem = os.path.join(d, S.token(cid))  # the oracle's to_token names the file
        assert np.array_equal(np.fromfile(stem + ".xyz", "<f8").reshape(-1, 3), P[idx])
        assert np.array_equal(np.fromfile(stem + ".rgb", np.uint8).reshape(-1, 3), rgb3[idx])
        assert np.array_equal(np.fromfile(stem + ".intensity", "<f4"), inten[idx])
    # S2Cells::from_data_provider: the loaded cloud answers like the built one
    back = ctx.load_s2_dir(d)
    assert back.num_points == n and np.array_equal(back.cell_ids, cloud.cell_ids) and np.array_equal(back.cell_counts, cloud.cell_counts)
    assert np.array_equal(back.bbox_min, cloud.bbox_min) and back.has_color and back.has_intensity
    centre = int(S.oracle_cell_ids(P[n // 2:n // 2 + 1], 20)[0])
    u = np.array([centre, S.orc().orc_s2_next(centre)], np.uint64)
    a, b = cloud.query_union(u), back.query_union(u)
    assert a["total"] == b["total"] > 0 and np.array_equal(a["xyz"], b["xyz"]) and np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["intensity"], b["intensity"])
    assert np.array_equal(back.query_union(None)["xyz"], P[want["order"].astype(np.int64)])
    back.free()
    cloud.free()
    # an octree directory is not an S2 cloud; a missing cell file is NodeNotFound
    o = Meta()
    o.version = 13
    o.octree.resolution = 1.0
    od = tmp_path / "oct"
    od.mkdir()
    (od / "meta.pb").write_bytes(o.SerializeToString())
    with pytest.raises(pcv._native.PcvError) as e:
        ctx.load_s2_dir(od)
    assert "does not describe S2 point clouds" in str(e.value)
    os.remove(os.path.join(d, pcv.s2_token(want["ids"][0]) + ".rgb"))
    with pytest.raises(pcv._native.PcvError) as e:
        ctx.load_s2_dir(d)
    assert e.value.code == -4


def test_split_two_million_points(ctx):
    """More points than one wave of the grid-stride kernels covers: cells, counts, per-cell order and a union query == oracle."""
    import point_cloud_viewer_b200 as pcv

    n = 2_000_000
    x, y, z, rgb = pcv.synth_points_host(pcv.SYNTH_SLAB_ECEF, 80293751232, 0, n)
    P = np.stack([x, y, z], 1)
    want = S.split(P, 22)
    cloud = ctx.build_s2_cloud(x, y, z, rgb, None, split_level=22)
    assert np.array_equal(cloud.cell_ids, want["ids"]) and np.array_equal(cloud.cell_counts, want["counts"])
    allp = cloud.query_union(None)
    assert np.array_equal(allp["src"], want["order"]) and np.array_equal(allp["rgb"], rgb.reshape(-1, 3)[want["order"].astype(np.int64)])
    u = S.oracle_cell_ids(P[:5], 17)
    inside = S.union_test(S.normalize(u), S.oracle_cell_ids(P, 30))[0]
    keep = want["order"][inside[want["order"].astype(np.int64)]]
    got = cloud.query_union(u)
    assert got["total"] == len(keep) > 1000 and np.array_equal(got["src"], keep)
    assert np.array_equal(ctx.s2_cell_ids(x, y, z, 22), S.oracle_cell_ids(P, 22))
    st = cloud.build_stats()
    assert st["ms_device"] > 0 and st["kernel_launches"] >= 5 and st["algorithmic_bytes"] == 2 * n * 27
    cloud.free()
