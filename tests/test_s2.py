"""S2 cell arithmetic without a GPU.  The `s2` crate is not vendored and the reference holds no S2 golden
vector, so the oracle's restatement is checked against the structural invariants S2 publishes; the product's csrc/s2.h (a
different formulation of the same curve) must then equal the oracle bit for bit."""
import numpy as np
import pytest

import s2_api as S

L = S.orc


def _ecef(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    return v * rng.uniform(6352800.0, 6384400.0, (n, 1))


def test_face_cells_and_tokens():
    axes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], float)
    ids = S.oracle_cell_ids(axes, 0)
    assert [int(i) for i in ids] == [(f << 61) | (1 << 60) for f in range(6)]
    assert [S.token(i) for i in ids] == ["1", "3", "5", "7", "9", "b"]
    assert S.token(0) == "X"
    # the leaf cell of (1, 0, 0): (i, j) = (2^29, 2^29) is the first leaf of the face centre's third child
    assert int(S.oracle_cell_ids(axes[:1], 30)[0]) == 0x1000000000000001
    assert S.token(0x1000000000000001) == "1000000000000001"


def test_hierarchy():
    P = _ecef(2000, 1)
    leaf = S.oracle_cell_ids(P, 30)
    prev = leaf
    for level in range(29, -1, -1):
        cur = S.oracle_cell_ids(P, level)
        lsb = cur & (~cur + np.uint64(1))
        assert (lsb == np.uint64(1) << np.uint64(2 * (30 - level))).all()
        assert ((cur - (lsb - np.uint64(1)) <= prev) & (prev <= cur + (lsb - np.uint64(1)))).all()  # the parent's range holds the child
        assert all(L().orc_s2_parent(int(p), level) == int(c) for p, c in zip(prev[:50], cur[:50]))
        prev = cur


@pytest.mark.parametrize("level", [1, 2, 5, 11, 20, 30])
def test_hilbert_curve_is_continuous_inside_a_face(level):
    """Consecutive cells along the curve share an edge: their (i, j) differ by one cell size in exactly one coordinate.  This
    exercises both look-up tables (ij -> pos on the way in, pos -> ij on the way out) over whole 4-level chunks."""
    rng = np.random.default_rng(level)
    size = 1 << (30 - level)
    for _ in range(300):
        f = int(rng.integers(0, 6))
        i, j = (int(v) for v in rng.integers(0, 1 << 30, 2))
        c = L().orc_s2_parent(L().orc_s2_from_face_ij(f, i, j), level)
        n = L().orc_s2_next(c)
        if (n >> 61) != (c >> 61):
            continue  # last cell of the face
        f0, i0, j0 = S.face_ij(c)
        f1, i1, j1 = S.face_ij(n)
        a, b = abs((i0 & -size) - (i1 & -size)), abs((j0 & -size) - (j1 & -size))
        assert f0 == f1 == f and sorted((a, b)) == [0, size], (hex(c), hex(n))
        assert S.face_ij(L().orc_s2_from_face_ij(f, i, j)) == (f, i, j)  # the two tables are inverse to each other


def test_curve_is_continuous_across_faces_and_centres_round_trip():
    """The last cell of face f touches the first cell of face f + 1 (the face (u, v) frames are chosen for exactly that), and
    the centre of a cell maps back to the cell: xyz -> face / uv and face / uv -> xyz agree in every face's sign conventions."""
    for level in (3, 10, 18):
        ang = 4.0 / (1 << level)  # well above one cell diagonal, far below a face
        for f in range(6):
            last = L().orc_s2_parent(((f + 1) << 61) - 1, level)
            first = L().orc_s2_parent(((f + 1) % 6) << 61 | 1, level)
            a, b = S.centre(last), S.centre(first)
            a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
            assert np.arccos(np.clip(a @ b, -1, 1)) < ang, (level, f)
    rng = np.random.default_rng(3)
    for level in (0, 1, 7, 20, 30):
        cells = [L().orc_s2_parent(L().orc_s2_from_face_ij(int(rng.integers(0, 6)), *(int(v) for v in rng.integers(0, 1 << 30, 2))), level) for _ in range(200)]
        C = np.array([S.centre(c) for c in cells])
        assert [int(v) for v in S.oracle_cell_ids(C * 6.37e6 / np.linalg.norm(C, axis=1)[:, None], level)] == cells
        assert all(L().orc_s2_level(c) == level for c in cells)


def test_product_cell_ids_equal_the_oracle():
    P = np.concatenate([_ecef(200000, 5), _ecef(1000, 6) / 6.37e6, _ecef(1000, 7) * 1e-300,
                        np.array([[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1], [-1, -1, -1], [0, 0, 0], [1e308, 1e308, 1e308], [0.0, -0.0, 5.0]], float)])
    for level in (30, 20, 12, 0):
        got, valid = S.product_cell_ids(P, level)
        assert np.array_equal(got, S.oracle_cell_ids(P, level)), level
    with np.errstate(over="ignore"):
        r = np.sqrt(P[:, 0] * P[:, 0] + P[:, 1] * P[:, 1] + P[:, 2] * P[:, 2])
    assert np.array_equal(valid, ~((r > 6384400.0) | (r < 6352800.0))) and valid[:200000].all() and not valid[200000:].any()
    rng = np.random.default_rng(8)
    for _ in range(2000):
        f, i, j = int(rng.integers(0, 6)), *(int(v) for v in rng.integers(0, 1 << 30, 2))
        c = S.tb().tbs_from_face_ij(f, i, j)
        assert c == L().orc_s2_from_face_ij(f, i, j) and S.tb().tbs_is_valid(c) and S.tb().tbs_level(c) == 30
    for cid in (0, 1 << 60, 0x1000000000000001, 0x89c25a31c0000000, 0xb000000000000000):
        assert S.token(cid, product=True) == S.token(cid)
        back = S.C.c_uint64()
        assert S.tb().tbs_from_token(S.token(cid).encode(), S.C.byref(back)) == 0 and back.value == cid
    assert not S.tb().tbs_is_valid(0) and not S.tb().tbs_is_valid(0xc000000000000001) and not S.tb().tbs_is_valid(0x1000000000000002)


def test_union_normalize_contains_intersects():
    rng = np.random.default_rng(9)
    P = _ecef(3000, 10)
    leaves = S.oracle_cell_ids(P, 30)
    # a messy union: cells of mixed levels, duplicates, children next to their parents, four siblings
    base = [int(v) for v in S.oracle_cell_ids(P[:40], 12)] + [int(v) for v in S.oracle_cell_ids(P[:60], 16)] + [int(v) for v in S.oracle_cell_ids(P[40:50], 9)]
    parent = int(S.oracle_cell_ids(P[100:101], 10)[0])
    lsb = parent & -parent
    kids = [parent - lsb + (lsb >> 2) * (2 * k + 1) // 1 for k in range(4)]  # the four children: parent.range_min + (2k+1) * child_lsb
    kids = [(parent - (lsb - 1)) + (lsb >> 2) - 1 + k * (lsb >> 1) for k in range(4)]
    assert all(L().orc_s2_parent(k, 10) == parent and L().orc_s2_level(k) == 11 for k in kids)
    cu_in = np.array(base + base[:7] + kids, np.uint64)
    cu = S.normalize(cu_in)
    assert np.array_equal(cu, S.normalize(cu_in, product=True))
    assert (np.diff(cu.astype(object)) > 0).all() and parent in [int(v) for v in cu] and not any(k in [int(v) for v in cu] for k in kids)
    lo = cu - ((cu & (~cu + np.uint64(1))) - np.uint64(1))
    hi = cu + ((cu & (~cu + np.uint64(1))) - np.uint64(1))
    assert (lo[1:].astype(object) > hi[:-1].astype(object)).all()  # disjoint ranges
    # every input cell is covered by the normalised union; brute-force membership agrees with the binary search
    probe = np.concatenate([leaves, S.oracle_cell_ids(P, 14), S.oracle_cell_ids(P, 8), cu_in])
    c, i = S.union_test(cu, probe)
    plo = probe - ((probe & (~probe + np.uint64(1))) - np.uint64(1))
    phi = probe + ((probe & (~probe + np.uint64(1))) - np.uint64(1))
    want_c = np.array([bool(((lo <= p) & (p <= hi)).any()) for p in probe])
    want_i = np.array([bool(((lo <= b) & (a <= hi)).any()) for a, b in zip(plo, phi)])
    assert np.array_equal(c, want_c) and np.array_equal(i, want_i) and c[-len(cu_in):].all() and 0 < c[:3000].sum() < 3000
    c2, i2 = S.union_test(cu, probe, product=True)
    assert np.array_equal(c2, c) and np.array_equal(i2, i)
    # the reference's own query shape (point_cloud_test/src/queries.rs:49-53): a cell and its successor
    cell = int(S.oracle_cell_ids(P[:1], 20)[0])
    pair = S.normalize(np.array([cell, L().orc_s2_next(cell)], np.uint64))
    assert len(pair) in (1, 2) and S.union_test(pair, S.oracle_cell_ids(P[:1], 30))[0][0]


def test_split_restatement():
    P = _ecef(5000, 11)
    P[100:200] = P[100]  # one crowded cell
    r = S.split(P, 12)
    ids = S.oracle_cell_ids(P, 12)
    assert r["ok"] and np.array_equal(r["ids"], np.unique(ids)) and r["counts"].sum() == len(P)
    assert np.array_equal(r["bmin"], P.min(0)) and np.array_equal(r["bmax"], P.max(0))
    o = 0
    for cid, cnt in zip(r["ids"], r["counts"]):
        assert np.array_equal(r["order"][o:o + int(cnt)], np.nonzero(ids == cid)[0])  # input order inside every cell
        o += int(cnt)
    P[777] *= 1.01  # "is not a valid ECEF point"
    bad = S.split(P, 12)
    assert not bad["ok"] and bad["bad_index"] == 777


# ---- meta.pb of an S2 cloud: the product's writer / reader against python-protobuf (descriptor restating proto.proto:58-149) --
def _tbs_meta():
    L = S.tb()
    L.tbs_encode_s2_meta.restype = S.C.c_int64
    L.tbs_encode_s2_meta.argtypes = [S.C.c_void_p, S.C.c_void_p, S.C.c_void_p, S.C.c_uint64, S.C.c_int, S.C.c_int, S.C.c_void_p, S.C.c_uint64]
    L.tbs_decode_s2_meta.restype = S.C.c_int64
    L.tbs_decode_s2_meta.argtypes = [S.C.c_void_p, S.C.c_uint64, S.C.c_void_p, S.C.c_void_p, S.C.c_void_p, S.C.c_uint64, S.C.POINTER(S.C.c_int), S.C.POINTER(S.C.c_int),
                                     S.C.POINTER(S.C.c_int), S.C.c_char_p, S.C.c_int]
    return L


def _decode(buf, cap=8192):
    L = _tbs_meta()
    bb = np.zeros(6)
    ids, counts = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    hc, hi, ver = S.C.c_int(), S.C.c_int(), S.C.c_int()
    err = S.C.create_string_buffer(256)
    raw = np.frombuffer(buf, np.uint8)
    n = L.tbs_decode_s2_meta(raw.ctypes.data, len(raw), bb.ctypes.data, ids.ctypes.data, counts.ctypes.data, cap, S.C.byref(hc), S.C.byref(hi), S.C.byref(ver), err, 256)
    return n, bb, ids[:max(n, 0)], counts[:max(n, 0)], bool(hc.value), bool(hi.value), ver.value, err.value.decode()


def test_s2_meta_against_python_protobuf():
    from proto_meta import Meta

    P = _ecef(4000, 21)
    r = S.split(P, 18)
    bb = np.concatenate([r["bmin"], r["bmax"]])
    L = _tbs_meta()
    for hc, hi in ((1, 1), (1, 0), (0, 0)):
        out = np.zeros(1 << 20, np.uint8)
        n = L.tbs_encode_s2_meta(bb.ctypes.data, r["ids"].ctypes.data, r["counts"].ctypes.data, len(r["ids"]), hc, hi, out.ctypes.data, len(out))
        assert n > 0
        raw = out[:n].tobytes()
        m = Meta.FromString(raw)  # an independent parser
        assert m.version == 13 and m.WhichOneof("data") == "s2" and not m.HasField("octree")
        assert [c.id for c in m.s2.cells] == [int(v) for v in r["ids"]] and [c.num_points for c in m.s2.cells] == [int(v) for v in r["counts"]]
        assert [(a.name, a.data_type) for a in m.s2.attributes] == ([("color", 27)] if hc else []) + ([("intensity", 11)] if hi else [])
        bbm = m.bounding_box
        assert [bbm.min.x, bbm.min.y, bbm.min.z, bbm.max.x, bbm.max.y, bbm.max.z] == list(bb)
        assert m.SerializeToString() == raw  # canonical field order, nothing unknown
        k, bb2, ids, counts, c2, i2, ver, err = _decode(raw)
        assert (k, ver, err, c2, i2) == (len(r["ids"]), 13, "", bool(hc), bool(hi)) and np.array_equal(ids, r["ids"]) and np.array_equal(counts, r["counts"]) and np.array_equal(bb2, bb)
    # what python-protobuf serialises (cells in another order, attributes swapped) is read back; other metas are rejected
    m = Meta()
    m.version = 13
    m.bounding_box.min.x, m.bounding_box.max.z = -1.5, 7.25
    for cid, cnt in ((0x89c25a31c0000000, 7), (0x1000000000000001, 1)):
        c = m.s2.cells.add()
        c.id, c.num_points = cid, cnt
    a = m.s2.attributes.add()
    a.name, a.data_type = "intensity", 11
    a = m.s2.attributes.add()
    a.name, a.data_type = "color", 27
    k, bb2, ids, counts, c2, i2, ver, err = _decode(m.SerializeToString())
    assert k == 2 and [int(v) for v in ids] == [0x89c25a31c0000000, 0x1000000000000001] and [int(v) for v in counts] == [7, 1] and c2 and i2 and bb2[0] == -1.5 and bb2[5] == 7.25
    m.version = 11
    assert _decode(m.SerializeToString())[7] == "No S2 point cloud supported with version 11"
    o = Meta()
    o.version = 13
    o.octree.resolution = 0.5
    assert _decode(o.SerializeToString())[7] == "This meta does not describe S2 point clouds"
    assert _decode(b"\\x0a\\xff\\xff")[0] == -1


def test_fuzz_product_against_oracle():
    """hypothesis: arbitrary finite vectors (any magnitude, signs, exact ties between components) - csrc/s2.h == oracle."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    comp = st.one_of(st.floats(allow_nan=False, allow_infinity=False, width=64), st.sampled_from([0.0, -0.0, 1.0, -1.0, 0.5, 6.37e6, -6.37e6, 1e-310, 1e300]))

    @hyp.settings(max_examples=300, deadline=None)
    @hyp.given(st.lists(st.tuples(comp, comp, comp), min_size=1, max_size=40), st.integers(0, 30))
    def run(pts, level):
        P = np.array(pts, np.float64)
        with np.errstate(all="ignore"):
            got, _ = S.product_cell_ids(P, level)
            assert np.array_equal(got, S.oracle_cell_ids(P, level))

    run()
