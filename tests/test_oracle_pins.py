"""Pin the oracle (CPU restatement) against every golden vector / known-answer test the reference holds for the
hot path.  No GPU needed."""
import ctypes as C
import math

import numpy as np

import oracle_api as O

L = O.lib()


def _cube(name, root_min, root_edge):
    hi, lo = O.id_from_str(name)
    out = (C.c_double * 4)()
    L.orc_find_bounding_cube(hi, lo, O._d(root_min), float(root_edge), out)
    return tuple(out)


# ---- src/octree/node.rs:277-317 -------------------------------------------------------------------------
def test_parent_node_name():
    hi, lo = O.id_from_str("r123456")
    phi, plo, ci = C.c_uint64(), C.c_uint64(), C.c_int()
    L.orc_node_id_parent(hi, lo, C.byref(phi), C.byref(plo), C.byref(ci))
    assert O.id_str(phi.value, plo.value) == "r12345"


def test_child_index():
    for name, want in (("r123451", 1), ("r123457", 7), ("r", -1)):
        hi, lo = O.id_from_str(name)
        phi, plo, ci = C.c_uint64(), C.c_uint64(), C.c_int()
        L.orc_node_id_parent(hi, lo, C.byref(phi), C.byref(plo), C.byref(ci))
        assert ci.value == want


def test_bounding_box_of_node_ids():
    assert _cube("r0", (-5, -5, -5), 10) == (-5.0, -5.0, -5.0, 5.0)
    assert _cube("r13", (-5, -5, -5), 10) == (-5.0, -2.5, 2.5, 2.5)


def test_node_id_round_trip_and_proto_halves():
    rng = np.random.default_rng(0)
    for _ in range(200):
        level = int(rng.integers(0, 41))
        name = "r" + "".join(str(int(d)) for d in rng.integers(0, 8, level))
        hi, lo = O.id_from_str(name)
        assert O.id_str(hi, lo) == name
        v = (hi << 64) | lo
        assert v >> 120 == level and (v & ((1 << 120) - 1)) == (int(name[1:], 8) if level else 0)
        if level < 40:
            chi, clo = C.c_uint64(), C.c_uint64()
            L.orc_node_id_child(hi, lo, 5, C.byref(chi), C.byref(clo))
            assert O.id_str(chi.value, clo.value) == name + "5"


# ---- src/read_write/codec.rs:154-212 (tolerances) + truncation behaviour ------------------------------------
def test_codec_round_trip_tolerances():
    value, mn, edge = 41.33333, 40.0, 2.0
    for enc, tol in ((3, 1e-7), (4, 1e-14), (1, 1e-2), (2, 1e-4)):
        code = L.orc_encode(value, mn, edge, enc)
        assert abs(L.orc_decode(code, mn, edge, enc) - value) < tol, enc


def test_codec_truncates_and_clamps():
    assert L.orc_encode(40.0 + 2.0 * 0.999999, 40.0, 2.0, 1) == 254  # trunc(255 * 0.999999)
    assert L.orc_encode(39.0, 40.0, 2.0, 2) == 0 and L.orc_encode(43.0, 40.0, 2.0, 2) == 65535
    assert L.orc_encode(float("nan"), 40.0, 2.0, 1) == 0  # NaN passes clamp, `as u8` -> 0


def test_position_encoding_thresholds():
    # bits = floor(log2(edge/res)) + 1 -> <=8 U8, <=16 U16, <=24 F32, else F64 (codec.rs:31-40)
    assert L.orc_position_encoding(200.0, 1.0) == 1  # the reference scenario: floor(log2 200)+1 = 8
    assert L.orc_position_encoding(256.0, 1.0) == 2
    assert L.orc_position_encoding(65535.0, 1.0) == 2 and L.orc_position_encoding(65536.0, 1.0) == 3
    assert L.orc_position_encoding(2.0 ** 24 - 1, 1.0) == 3 and L.orc_position_encoding(2.0 ** 24, 1.0) == 4
    assert L.orc_position_encoding(0.5, 1.0) == 1  # negative log2 saturates to 0


# ---- src/math/sat.rs:214-268 ---------------------------------------------------------------------------------
def _cube_isec(lo, hi):
    c = []
    for x in (lo, hi):
        for y in (lo, hi):
            for z in (lo, hi):
                c += [x, y, z]
    return O._d(c)


def test_sat_cube_with_cube():
    unit = O._d([1, 0, 0, 0, 1, 0, 0, 0, 1])
    c1, c2, c3 = _cube_isec(-1.0, 1.0), _cube_isec(-0.5, 1.5), _cube_isec(-0.9, -0.7)

    def isect(a, b):
        return L.orc_intersector_intersect(a, unit, 3, unit, 3, b, unit, 3, unit, 3)

    IN, CROSS, OUT = 0, 1, 2
    assert isect(c1, c2) == CROSS
    assert isect(c2, c3) == OUT
    assert isect(c1, c3) == IN
    assert isect(c3, c1) == CROSS


# ---- src/geometry/obb.rs:100-140 ------------------------------------------------------------------------------
def _obb_loc(quat, half):
    import point_cloud_viewer_b200.geometry as G

    loc = G.obb(G.Isometry((0, 0, 0), quat), half)
    o = O.Location()
    for f, _ in O.Location._fields_:
        setattr(o, f, getattr(loc, f))
    return o


def test_obb_intersects_aabb():
    import point_cloud_viewer_b200.geometry as G

    half = (1.0, 2.0, 3.0)
    bmin, bmax = O._d([0.5, 1.0, -3.0]), O._d([1.5, 3.0, 3.0])
    axes = (C.c_double * (3 * 64))()
    zero = _obb_loc((0, 0, 0, 1), half)
    assert L.orc_cached_axes(C.byref(zero), axes, 64) == 3
    assert L.orc_cached_intersect_aabb(C.byref(zero), bmin, bmax) == 1  # Cross
    r45 = _obb_loc(G.quat_from_axis_angle([0, 0, 1], math.pi / 4.0), half)
    assert L.orc_cached_axes(C.byref(r45), axes, 64) == 5
    assert L.orc_cached_intersect_aabb(C.byref(r45), bmin, bmax) == 2  # Out
    arb = _obb_loc(G.quat_from_axis_angle([0.2, 0.5, -0.7], 0.123), half)
    assert L.orc_cached_axes(C.byref(arb), axes, 64) == 15


# ---- src/math/mod.rs:192-220, src/geometry/frustum.rs:178-205 ---------------------------------------------------
def test_perspective_inverse_and_fov():
    import point_cloud_viewer_b200.geometry as G

    p = G.Perspective(-0.123, 0.45, 0.04, 0.75, 1.0, 4.0)
    ref = (C.c_double * 16)()
    assert L.orc_try_inverse(O._d(p.matrix.T.reshape(-1)), ref) == 1
    assert np.abs(np.array(ref).reshape(4, 4).T - p.inverse()).max() < 1e-6
    # Perspective::new_fov vs nalgebra::Perspective3::new(aspect, fovy, near, far) (frustum.rs:185-205)
    a = G.Perspective.new_fov(1.2, 0.66, 1.0, 100.0).matrix
    t = math.tan(0.66 / 2.0)
    b = np.zeros((4, 4))
    b[0, 0], b[1, 1] = 1.0 / (1.2 * t), 1.0 / t
    b[2, 2], b[2, 3], b[3, 2] = (100.0 + 1.0) / (1.0 - 100.0), 2.0 * 100.0 * 1.0 / (1.0 - 100.0), -1.0
    assert np.allclose(a, b, rtol=1e-15, atol=0)


def test_frustum_intersects_aabb():
    import point_cloud_viewer_b200.geometry as G

    rot = G.Isometry((0, 0, 0), G.quat_from_axis_angle([1, 0, 0], math.pi))
    loc = G.frustum(rot, G.Perspective(-0.5, 0.0, -0.5, 0.0, 1.0, 4.0))
    o = O.Location()
    for f, _ in O.Location._fields_:
        setattr(o, f, getattr(loc, f))
    bmin, bmax = (-0.5, 0.25, 1.5), (-0.25, 0.5, 3.5)
    assert L.orc_location_intersect_aabb_generic(C.byref(o), O._d(bmin), O._d(bmax)) == 0  # Relation::In
    assert L.orc_location_contains(C.byref(o), O._d(bmin)) == 1
    assert L.orc_location_contains(C.byref(o), O._d(bmax)) == 1


# ---- src/octree/tests.rs + hand-derived golden tree ---------------------------------------------------
def _scenario():
    n = 100001
    x, y, z = np.zeros(n), np.zeros(n), np.zeros(n)
    x[-1], y[-1], z[-1] = -200.0, -40.0, 30.0
    rgb = np.tile(np.array([255, 0, 0], np.uint8), (n, 1))
    return O.build(x, y, z, rgb, 1.0, (-200, -40, 0), (0, 0, 30)), n


def test_reference_scenario_golden_tree():
    t, n = _scenario()
    assert {k: (v["num_points"], v["enc"]) for k, v in t.nodes.items()} == {"r": (12501, 1), "r0": (0, 1), "r4": (87500, 1)}
    assert t.nodes["r"]["cube"] == (-200.0, -40.0, 0.0, 200.0) and t.nodes["r4"]["cube"] == (-100.0, -40.0, 0.0, 100.0)
    _, _, _, src = t.node_data("r")
    assert src[0] == n - 1 and np.array_equal(src[1:], np.arange(0, 100000, 8))  # outlier (child 0) first, then r4[0::8]
    _, _, _, src4 = t.node_data("r4")
    assert np.array_equal(src4, np.array([i for i in range(100000) if i % 8]))
    assert sum(v["num_points"] for v in t.nodes.values()) == n


def test_reference_scenario_all_points_query_and_heap():
    t, n = _scenario()
    loc = O.Location()
    loc.kind = 0
    assert t.nodes_in_location(loc) == ["r", "r0", "r4"]  # zero-point nodes are still visited (no num_points filter)
    q = t.query(loc)
    assert len(q["src"]) == n and q["tested"] == n


def test_point_culling_equals_sat_with_face_normals():
    """point_cloud_test/tests/main.rs:104-127: contains() == SAT over the face normals against a single point."""
    import point_cloud_viewer_b200.geometry as G

    rng = np.random.default_rng(3)
    iso = G.Isometry((5.0, -3.0, 2.0), G.quat_from_axis_angle([0.3, -0.2, 0.9], 0.8))
    locs = [G.aabb((-20, -20, -5), (30, 25, 6)), G.obb(iso, (25.0, 25.0, 5.0)), G.frustum(iso, G.Perspective.new_fov(1.0, 1.2, 0.1, 60.0))]
    pts = rng.random((3000, 3)) * [200, 200, 40] - [100, 100, 20]
    hits = 0
    for loc in locs:
        o = O.Location()
        for f, _ in O.Location._fields_:
            setattr(o, f, getattr(loc, f))
        for p in pts:
            a = L.orc_location_contains(C.byref(o), O._d(p))
            b = L.orc_location_contains_sat(C.byref(o), O._d(p))
            if loc.kind == 1:
                # Aabb::contains is half-open [min,max) while SAT is closed: they can only differ on the max faces
                assert a == b or np.any(p == np.array(loc.aabb_max))
            else:
                assert a == b
            hits += a
    assert hits > 0
