"""GPU parity of the query side (through the C ABI) against the oracle: node selection (SAT BFS), filtered point
streaming, LOD visible-node order, X-ray tiles, on-disk round trip, batch delivery semantics."""
import math
import os

import numpy as np
import pytest

import oracle_api as O
from parity import compare_trees

pytestmark = pytest.mark.gpu


def _copy_loc(loc):
    o = O.Location()
    for f, _ in O.Location._fields_:
        setattr(o, f, getattr(loc, f))
    return o


@pytest.fixture(scope="module")
def scene():
    """200k-point slab in the ECEF-like frame with intensity, small max_points_per_node -> a deep tree; built by
    both the CUDA path and the oracle (and checked identical first)."""
    import point_cloud_viewer_b200 as pcv

    n = 200_000
    x, y, z, rgb = pcv.synth_points_host(pcv.SYNTH_SLAB_ECEF, 80293751232, 0, n)
    inten = (np.arange(n) % 1000).astype(np.float32)
    bmin, bmax, res = pcv.synth_bbox(pcv.SYNTH_SLAB_ECEF)
    c = pcv.Context(0, max_points_per_node=3000)
    tree = c.build_octree(x, y, z, rgb, res, bmin, bmax, intensity=inten)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, intensity=inten, max_points_per_node=3000)
    compare_trees(ref, tree)
    yield dict(pcv=pcv, ctx=c, tree=tree, ref=ref, P=np.stack([x, y, z], 1), bmin=bmin, bmax=bmax, n=n)
    tree.free()
    c.close()


def _locations(s):
    """The query shapes of point_cloud_test/src/queries.rs (Aabb 0.2..0.8 of the bbox, centred OBB with half the
    data's half-extent, Perspective3(1.0, 1.2, 0.1, 10.0) frustum at the slab pose) + extra cases."""
    pcv = s["pcv"]
    G = pcv.geometry
    bmin, bmax = s["bmin"], s["bmax"]
    d = bmax - bmin
    # local frame of the synthetic slab: rotation Rz(0.7)*Ry(-0.9), translation as in csrc/synth.cuh
    q = G.quat_mul(G.quat_from_axis_angle([0, 0, 1], 0.7), G.quat_from_axis_angle([0, 1, 0], -0.9))
    ecef_from_local = G.Isometry((4157222.543, 664789.307, 4774952.099), q)
    locs = {
        "all": G.all_points(),
        "aabb": G.aabb(bmin + 0.2 * d, bmin + 0.8 * d),
        "aabb_small": G.aabb(bmin + 0.45 * d, bmin + 0.5 * d),
        "obb": G.obb(ecef_from_local, (50.0, 50.0, 5.0)),
        "frustum": G.frustum(ecef_from_local, G.Perspective.new_fov(1.0, 1.2, 0.1, 10.0)),
        "frustum_far": G.frustum(ecef_from_local * G.Isometry((0, 0, 0), G.quat_from_axis_angle([1, 0.3, 0], 1.3)), G.Perspective.new_fov(1.3, 0.9, 0.5, 150.0)),
        "obb_tilted": G.obb(ecef_from_local * G.Isometry((10, -20, 1), G.quat_from_axis_angle([0.2, 0.5, -0.7], 0.523)), (30.0, 12.0, 4.0)),
    }
    return locs


def test_nodes_in_location(scene):
    for name, loc in _locations(scene).items():
        got = scene["tree"].nodes_in_location(loc)
        want = scene["ref"].nodes_in_location(_copy_loc(loc))
        assert got == want, name
        assert len(got) > 0, name


def test_query_points_match_oracle_and_brute_force(scene):
    tree, ref, P = scene["tree"], scene["ref"], scene["P"]
    for name, loc in _locations(scene).items():
        want = ref.query(_copy_loc(loc), with_intensity=True)
        batches = tree.query_points(loc, batch_size=7777)
        assert all(len(b["src"]) == 7777 for b in batches[:-1]), name
        src = np.concatenate([b["src"] for b in batches]) if batches else np.zeros(0, np.uint64)
        xyz = np.concatenate([b["xyz"] for b in batches]) if batches else np.zeros((0, 3))
        rgb = np.concatenate([b["rgb"] for b in batches]) if batches else np.zeros((0, 3), np.uint8)
        inten = np.concatenate([b["intensity"] for b in batches]) if batches else np.zeros(0, np.float32)
        # same nodes in the same (BFS) order, file order inside a node -> identical streams
        assert np.array_equal(src, want["src"]), name
        assert np.array_equal(xyz, want["xyz"]), name  # decoded f64 coordinates bit-equal (north star asks 1e-6 relative)
        assert np.array_equal(rgb, want["rgb"]), name
        assert np.array_equal(inten, want["intensity"]), name
        assert len(src) > 0, name
        # vs brute-force culling of the ORIGINAL points (tests/main.rs:199-202: <= 1 % index mismatches)
        keep = np.array([O.lib().orc_location_contains(_copy_loc(loc), O._d(p)) for p in P[:: max(1, len(P) // 20000)]], bool)
        sub = np.arange(len(P))[:: max(1, len(P) // 20000)]
        got = np.isin(sub, src.astype(np.int64))
        assert (got != keep).sum() <= max(1, math.ceil(min(got.sum(), keep.sum()) / 100) + 2), name


def test_query_interval_filter(scene):
    loc = _locations(scene)["aabb"]
    want = scene["ref"].query(_copy_loc(loc), filters=[100.0, 250.0], with_intensity=True)
    got = scene["tree"].query_points(loc, filters=[100.0, 250.0], batch_size=1 << 20)
    src = np.concatenate([b["src"] for b in got])
    assert np.array_equal(src, want["src"]) and len(src) > 0
    assert all(((b["intensity"] >= 100) & (b["intensity"] <= 250)).all() for b in got)


def test_query_batch_device_counts(scene):
    locs = list(_locations(scene).values())
    counts, tested = scene["tree"].query_batch_device(locs)
    for i, loc in enumerate(locs):
        want = scene["ref"].query(_copy_loc(loc))
        assert counts[i] == len(want["src"]) and tested[i] == want["tested"], i


def test_batch_iterator_semantics():
    """src/octree/tests.rs:83-136 on the 100 001-point octree: batch 5000, consumer errors at >= 13 000 points ->
    exactly 3 callbacks of 5000; batch N/2 without error -> exactly N points."""
    import point_cloud_viewer_b200 as pcv

    n = 100001
    x, y, z = np.zeros(n), np.zeros(n), np.zeros(n)
    x[-1], y[-1], z[-1] = -200.0, -40.0, 30.0
    rgb = np.tile(np.array([255, 0, 0], np.uint8), n)
    c = pcv.Context(0)
    tree = c.build_octree(x, y, z, rgb, 1.0, (0, 0, 0), (-200, -40, 30))
    state = dict(points=0, calls=0)

    def consume(b):
        state["calls"] += 1
        state["points"] += len(b["src"])
        return state["points"] >= 13000

    with pytest.raises(pcv.PcvError) as e:
        tree.query_points(pcv.geometry.all_points(), callback=consume, batch_size=5000)
    assert e.value.code == -5 and state["calls"] == 3 and state["points"] == 15000
    got = tree.query_points(pcv.geometry.all_points(), batch_size=n // 2)
    assert sum(len(b["src"]) for b in got) == n
    tree.free()
    c.close()


def _cameras(scene, k):
    G = scene["pcv"].geometry
    rng = np.random.default_rng(11)
    out = []
    for i in range(k):
        eye = scene["bmin"] + rng.random(3) * (scene["bmax"] - scene["bmin"])
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        iso = G.Isometry(eye, q)
        persp = G.Perspective.new_fov(1.0 + rng.random(), 0.5 + rng.random(), 0.1, 10.0 if i % 2 else 300.0)
        out.append(persp.matrix @ iso.inverse().to_homogeneous())
    return out


def test_get_visible_nodes(scene):
    some = 0
    for M in _cameras(scene, 12):
        got = scene["tree"].get_visible_nodes(M)
        want = scene["ref"].visible_nodes(M.T.reshape(-1))
        assert got == want
        some += len(got)
    assert some > 0
    with pytest.raises(scene["pcv"].PcvError) as e:
        scene["tree"].get_visible_nodes(np.zeros((4, 4)))
    assert e.value.code == -7


TRANSPARENT = np.array([255, 255, 255, 0], np.uint8)  # TRANSPARENT.to_u8(), src/color.rs:154-159: pixels without points


def test_xray_tiles(scene):
    tree, ref = scene["tree"], scene["ref"]
    bmin, bmax = scene["bmin"], scene["bmax"]
    d = bmax - bmin
    tmin, tmax = bmin + [0.3, 0.3, 0.0] * d, bmin + [0.6, 0.6, 1.0] * d
    any_g, rgba, zb = tree.xray_tile(tmin, tmax, 64, 48, want_bits=True)
    any_o, rgba_o, zb_o, zover_o = ref.xray_tile(tmin, tmax, 64, 48)
    assert any_g and any_o
    assert np.array_equal(zb, zb_o) and np.array_equal(rgba, rgba_o)
    assert (rgba[..., 3] == 255).any() and (rgba[..., 3] == 0).any()
    # with a query frame (OBB location + isometry transform of every point, generation.rs:471-497)
    G = scene["pcv"].geometry
    q = G.quat_mul(G.quat_from_axis_angle([0, 0, 1], 0.7), G.quat_from_axis_angle([0, 1, 0], -0.9))
    query_from_global = G.Isometry((4157222.543, 664789.307, 4774952.099), q).inverse()
    qmin, qmax = np.array([-40.0, -30.0, -10.0]), np.array([24.0, 34.0, 10.0])
    any_g, rgba, zb = tree.xray_tile(qmin, qmax, 128, 128, query_from_global=query_from_global.as7(), want_bits=True)
    any_o, rgba_o, zb_o, zover_o = ref.xray_tile(qmin, qmax, 128, 128, query_from_global=query_from_global.as7())
    assert any_g and any_o and np.array_equal(zb, zb_o) and np.array_equal(rgba, rgba_o)
    assert (rgba[..., 3] == 255).sum() > 1000
    # empty tile -> None in the reference
    any_g, rgba, _ = tree.xray_tile(bmax + 10, bmax + 20, 8, 8)
    assert not any_g and (rgba == TRANSPARENT).all()  # the reference returns None; the buffer holds only background


def test_xray_other_colouring_strategies(scene):
    """xray/src/generation.rs:200-405 with Binning = None: point colour mean, intensity mean, height stddev through Jet /
    Purplish.  The reference accumulates in (unspecified) arrival order in f32 / Welford-f64, so agreement is defined up to
    rounding: the set of covered pixels must be identical and every channel within one grey level of the oracle."""
    pcv = scene["pcv"]
    tree, ref = scene["tree"], scene["ref"]
    bmin, bmax = scene["bmin"], scene["bmax"]
    d = bmax - bmin
    tmin, tmax = bmin + [0.2, 0.2, 0.0] * d, bmin + [0.7, 0.7, 1.0] * d
    G = pcv.geometry
    q = G.quat_mul(G.quat_from_axis_angle([0, 0, 1], 0.7), G.quat_from_axis_angle([0, 1, 0], -0.9))
    qfg = G.Isometry((4157222.543, 664789.307, 4774952.099), q).inverse().as7()
    qmin, qmax = np.array([-40.0, -30.0, -10.0]), np.array([24.0, 34.0, 10.0])
    cases = [(pcv.XRAY_COLORED, 0.0, 0.0, 0), (pcv.XRAY_HEIGHT_STDDEV, 0.8, 0.0, 0), (pcv.XRAY_HEIGHT_STDDEV, 2.5, 0.0, 1)]
    if tree.has_intensity:
        cases.append((pcv.XRAY_INTENSITY, 0.0, 1000.0, 0))
        cases.append((pcv.XRAY_INTENSITY, 100.0, 800.0, 0))
    for mode, p0, p1, cm in cases:
        for (lo, hi, w, h, frame) in ((tmin, tmax, 96, 64, None), (qmin, qmax, 128, 128, qfg)):
            any_g, got = tree.xray_tile_attr(lo, hi, w, h, mode, p0, p1, cm, query_from_global=frame)
            any_o, want = ref.xray_tile_attr(lo, hi, w, h, mode, p0, p1, cm, query_from_global=frame)
            assert any_g and any_o
            assert np.array_equal(got[..., 3], want[..., 3]), (mode, "covered pixels")
            diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
            assert diff.max() <= 1, (mode, p0, p1, cm, int(diff.max()), int((diff > 0).sum()))
            assert (diff > 0).mean() < 0.02
            assert (got[..., 3] == 255).sum() > 200 and got[..., :3].std() > 0
    any_g, rgba = tree.xray_tile_attr(bmax + 10, bmax + 20, 8, 8, pcv.XRAY_COLORED)
    assert not any_g and (rgba == TRANSPARENT).all()


def test_on_disk_round_trip(scene, tmp_path):
    """a8-a10: <dir>/<id>.xyz|.rgb|.intensity + meta.pb; the oracle's reader loads what the product wrote and vice versa."""
    tree, ref = scene["tree"], scene["ref"]
    d1, d2 = str(tmp_path / "gpu"), str(tmp_path / "oracle")
    os.makedirs(d2)
    tree.write_dir(d1)
    ref.write_dir(d2)
    files1, files2 = sorted(os.listdir(d1)), sorted(os.listdir(d2))
    assert files1 == files2
    for f in files1:
        if f != "meta.pb":  # meta.pb node order is unspecified in the reference (FnvHashMap iteration)
            assert open(os.path.join(d1, f), "rb").read() == open(os.path.join(d2, f), "rb").read(), f
    assert not os.path.exists(os.path.join(d1, "r0.xyz")) or tree.nodes["r0"]["num_points"] > 0
    back = O.load_dir(d1)  # oracle reads the product's directory
    assert {k: (v["num_points"], v["enc"], v["cube"]) for k, v in back.nodes.items()} == {k: (v["num_points"], v["enc"], v["cube"]) for k, v in ref.nodes.items()}
    again = scene["ctx"].load_dir(d2)  # product reads the oracle's directory
    assert {k: (v["num_points"], v["enc"], v["cube"]) for k, v in again.nodes.items()} == {k: (v["num_points"], v["enc"], v["cube"]) for k, v in tree.nodes.items()}
    for name in list(tree.nodes)[:50]:
        if tree.nodes[name]["num_points"]:
            a, b = again.node_data(name), tree.node_data(name)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    loc = _locations(scene)["frustum_far"]
    assert again.nodes_in_location(loc) == tree.nodes_in_location(loc)
    again.free()
