"""The X-ray pipeline beyond the leaf tile without a GPU: the oracle's restatement against the reference's
known-answer vectors and independent witnesses, and the product's per-element code (csrc/xray_pyramid.h, through the
sequential test drivers) against the oracle."""
import numpy as np
import pytest

import oracle_api as O
import tbx_api as T
import point_cloud_viewer_b200 as pcv


def _smooth(n, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n] / n
    img = np.zeros((n, n, 4), np.float64)
    for k in range(4):
        a, b, c = rng.uniform(0.5, 3.0, 3)
        img[..., k] = 127.5 + 120 * np.sin(a * xx * 6 + c) * np.cos(b * yy * 6)
    return np.clip(img, 0, 255).astype(np.uint8)


def _noise(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, n, 4), dtype=np.uint8)


# ---- quadtree ids and rectangles: the reference's own vectors (quadtree/src/lib.rs:357-405) -----------------------
def test_quadtree_node_ids_reference_vectors():
    assert pcv.xray_node_id("r123210") == (6, int("123210", 4))
    lvl, idx = pcv.xray_node_id("r123210")
    assert pcv.xray_node_name(lvl - 1, idx >> 2) == "r12321"  # test_parent_node_name
    assert pcv.xray_node_id("r123321")[1] & 3 == 1 and pcv.xray_node_id("r123323")[1] & 3 == 3  # test_child_index
    for name in ("r", "r0", "r123323"):  # test_to_string
        assert pcv.xray_node_name(*pcv.xray_node_id(name)) == name
    # test_spatial_node_id_from_node_id: r301 is the cell (x = 4, y = 5) of the 8 x 8 grid at level 3; Node::get_child moves
    # +x for bit 1 and +y for bit 0, so with a root rect (0, 0, 8) its rect is min (4, 5), edge 1
    assert T.quad_rect_of(3, int("301", 4), (0.0, 0.0, 8.0)) == (4.0, 5.0, 1.0)
    assert T.quad_rect_of(0, 0, (1.5, -2.5, 8.0)) == (1.5, -2.5, 8.0)


def test_rect_and_levels():
    # find_quadtree_bounding_rect_and_levels (generation.rs:515-533): doubling from tile_size_px * pixel_size_m
    assert T.rect_and_levels((0, 0, 0), (1024, 1024, 1024), 4096, 0.0625) == ((0.0, 0.0, 1024.0), 2)  # BASELINE config 5: 16 leaf tiles
    assert T.rect_and_levels((10, 20, 0), (11, 21, 5), 256, 0.01) == ((10.0, 20.0, 2.56), 0)
    assert T.rect_and_levels((10, 20, 0), (13, 21, 5), 256, 0.01) == ((10.0, 20.0, 5.12), 1)
    assert T.rect_and_levels((0, 0, 0), (1, 1, 1), 256, 0.0) is None and T.rect_and_levels((0, 0, 0), (1, 1, 1), 256, float("nan")) is None


# ---- Lanczos3 reduction -------------------------------------------------------------------------------------------------
def _lanczos_f64(img, nw, nh):
    """Independent float64 evaluation of the same formula (separable, vertical then horizontal, u8 intermediate)."""
    def taps(o, n_in, n_out):
        ratio = n_in / n_out
        sr = max(ratio, 1.0)
        centre = (o + 0.5) * ratio
        left = min(max(int(np.floor(centre - 3 * sr)), 0), n_in - 1)
        right = min(max(int(np.ceil(centre + 3 * sr)), left + 1), n_in)
        x = (np.arange(left, right) - (centre - 0.5)) / sr
        w = np.where(np.abs(x) < 3, np.sinc(x) * np.sinc(x / 3), 0.0)
        return left, w

    h, w_, _ = img.shape
    tmp = np.zeros((nh, w_, 4))
    for oy in range(nh):
        l, w = taps(oy, h, nh)
        tmp[oy] = np.tensordot(w, img[l:l + len(w)].astype(np.float64), 1) / w.sum()
    tmp_q = np.clip(tmp, 0, 255)
    tmp8 = np.floor(tmp_q + 0.5)
    out = np.zeros((nh, nw, 4))
    for ox in range(nw):
        l, w = taps(ox, w_, nw)
        out[:, ox] = np.tensordot(tmp8[:, l:l + len(w)], w, ([1], [0])) / w.sum()
    return tmp_q, np.clip(out, 0, 255)


def test_resize_restatement_matches_float64_evaluation():
    for n, m, seed in ((32, 16, 1), (64, 32, 2), (30, 20, 3)):
        img = _noise(n, seed)
        got = O.resize_lanczos3(img, m, m).astype(np.int32)
        _, want = _lanczos_f64(img, m, m)
        # f32 against f64 arithmetic only moves results that sit on a rounding boundary; the intermediate image can differ
        # by one there, which the horizontal pass spreads with weight < 1
        d = np.abs(got - np.floor(want + 0.5))
        assert d.max() <= 2 and (d > 0).mean() < 0.02, (n, m, d.max(), (d > 0).mean())
    flat = np.full((16, 16, 4), 200, np.uint8)  # a constant image stays constant (round to nearest)
    assert (O.resize_lanczos3(flat, 8, 8) == 200).all()


def test_resize_close_to_pillow_lanczos():
    """Pillow implements the same filter (support 3 x scale, normalised weights) in fixed point with rounding: an
    independent implementation, so closeness is what is expected."""
    Image = pytest.importorskip("PIL.Image")
    img = _smooth(64, 5)
    got = O.resize_lanczos3(img, 32, 32).astype(np.float64)
    # channel by channel: Pillow resamples RGBA with premultiplied alpha, image-rs treats the four channels alike
    ref = np.stack([np.asarray(Image.fromarray(np.ascontiguousarray(img[..., k]), "L").resize((32, 32), Image.LANCZOS)) for k in range(4)], -1).astype(np.float64)
    d = got - ref
    assert np.abs(d).max() <= 2 and abs(d.mean()) < 0.25, (np.abs(d).max(), d.mean())


# ---- product code (csrc/xray_pyramid.h) against the oracle ------------------------------------------------------------------
@pytest.mark.parametrize("child_px,tile_px,missing", [(16, 16, ()), (16, 16, (0, 3)), (33, 33, (1,)), (8, 12, (2,)), (20, 8, (0, 1, 2)), (64, 64, ())])
def test_parent_tile_equals_oracle(child_px, tile_px, missing):
    ch = [None if k in missing else (_noise(child_px, 10 + k) if k % 2 else _smooth(child_px, 20 + k)) for k in range(4)]
    for bg in ((255, 255, 255, 255), (255, 255, 255, 0), (12, 34, 56, 78)):
        want, mosaic = O.build_parent_tile(ch, bg, tile_px, want_mosaic=True)
        got = T.build_parent(ch, bg, tile_px)
        assert np.array_equal(got, want)
        # build_parent's layout (generation.rs:433-446): child 1 top left, 0 bottom left, 3 top right, 2 bottom right
        cs = child_px
        for k, (y0, x0) in {1: (0, 0), 0: (cs, 0), 3: (0, cs), 2: (cs, cs)}.items():
            blk = mosaic[y0:y0 + cs, x0:x0 + cs]
            assert np.array_equal(blk, ch[k]) if ch[k] is not None else (blk == np.array(bg, np.uint8)).all()


def test_background_equals_oracle():
    img = _noise(40, 3)
    img[::3, ::2, 3] = 127
    img[1::3, ::2, 3] = 128
    for bg in ((255, 255, 255, 255), (255, 255, 255, 0), (1, 2, 3, 4)):
        want = O.assign_background(img, bg)
        assert np.array_equal(T.background(img, bg), want)
        assert (want[img[..., 3] < 128] == np.array(bg, np.uint8)).all() and np.array_equal(want[img[..., 3] >= 128], img[img[..., 3] >= 128])


def test_bin_of_is_rusts_cast():
    # (e as f64 / size) as i64: truncation toward zero, saturation, NaN -> 0 (generation.rs:143-145)
    assert T.bin_of(7.9, 2.0) == 3 and T.bin_of(-7.9, 2.0) == -3 and T.bin_of(0.0, 5.0) == 0
    assert T.bin_of(float("nan"), 1.0) == 0
    assert T.bin_of(3e38, 1e-30) == 2 ** 63 - 1 and T.bin_of(-3e38, 1e-30) == -(2 ** 63)
    assert T.bin_of(float("inf"), 1.0) == 2 ** 63 - 1 and T.bin_of(float("-inf"), 1.0) == -(2 ** 63)
    assert T.bin_of(np.float32(0.1), 0.1) == int(float(np.float32(0.1)) / 0.1)


def _binned_reference(pixel, attr, value, bin_size, npix):
    cols = {}
    for p, a, v in zip(pixel, attr, value):
        key = (int(p), T.bin_of(a, bin_size))
        s, c = cols.get(key, (np.zeros(value.shape[1], np.float32), 0))
        cols[key] = (s + v.astype(np.float32), c + 1)
    pix_sum = np.zeros((npix, value.shape[1]), np.float64)
    bins = np.zeros(npix, np.uint32)
    for (p, _), (s, c) in cols.items():
        pix_sum[p] += (s / np.float32(c)).astype(np.float64)
        bins[p] += 1
    return pix_sum, bins


@pytest.mark.parametrize("ncomp,stride", [(3, 4), (1, 1)])
def test_binned_aggregation(ncomp, stride):
    rng = np.random.default_rng(7)
    n, npix = 20000, 257
    pixel = rng.integers(0, npix, n).astype(np.uint32)
    pixel[:500] = 3  # one crowded column
    attr = rng.uniform(-50, 1000, n).astype(np.float32)
    attr[::97] = np.float32("inf")
    attr[5::97] = np.float32("-inf")  # the bin i64::MIN: the dedicated id
    attr[7::97] = np.float32("nan")
    value = rng.uniform(0, 1, (n, ncomp)).astype(np.float32)
    err, pix_sum, bins = T.binned(pixel, attr, value, 12.5, npix, stride)
    want_sum, want_bins = _binned_reference(pixel, attr, value, 12.5, npix)
    assert err == 0 and np.array_equal(bins, want_bins)
    assert np.allclose(pix_sum[:, :ncomp], want_sum, rtol=1e-5, atol=1e-5)
    assert bins.max() > 10 and (pix_sum[:, ncomp:] == 0).all()
    # more distinct bins than the bin table holds -> the error flag, not a hang or a wrong picture
    err, _, _ = T.binned(pixel[:300], np.arange(300, dtype=np.float32), value[:300], 1.0, npix, stride, bin_cap=64)
    assert err == 1
    # a column table exactly as large as the number of distinct columns still terminates
    err, s2, b2 = T.binned(pixel[:50], attr[:50], value[:50], 12.5, npix, stride, col_cap=50)
    assert err == 0 and b2.sum() == len({(int(p), T.bin_of(a, 12.5)) for p, a in zip(pixel[:50], attr[:50])})


# ---- the oracle's binned tile and quadtree on a small octree ------------------------------------------------------------
@pytest.fixture(scope="module")
def small_tree():
    n = 60000
    x, y, z, rgb = pcv.synth_points_host(pcv.SYNTH_SLAB_ECEF, 80293751232, 0, n)
    bmin, bmax, res = pcv.synth_bbox(pcv.SYNTH_SLAB_ECEF)
    inten = (np.arange(n) % 1000).astype(np.float32)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, max_points_per_node=5000, intensity=inten)
    return ref, np.asarray(bmin), np.asarray(bmax)


def test_oracle_binned_tile_reduces_to_unbinned_for_one_bin(small_tree):
    ref, bmin, bmax = small_tree
    # a bin size larger than every intensity puts each column into the single bin 0: Binning = Some is then Binning = None
    for mode, p0, p1 in ((1, 0.0, 0.0), (2, 0.0, 2000.0)):
        a1, one = ref.xray_tile_attr_binned(bmin, bmax, 64, 64, mode, 1e9, p0, p1)
        a0, none = ref.xray_tile_attr(bmin, bmax, 64, 64, mode, p0, p1, 0)
        assert a1 and a0 and np.array_equal(one, none)
    _, fine = ref.xray_tile_attr_binned(bmin, bmax, 64, 64, 1, 10.0)
    assert np.array_equal(fine[..., 3], none[..., 3]) and not np.array_equal(fine, one)
    # pixels without points are TRANSPARENT.to_u8() = (255, 255, 255, 0) (color.rs:154-159), not zero
    empty = fine[fine[..., 3] == 0]
    assert len(empty) and (empty == np.array([255, 255, 255, 0], np.uint8)).all()


def test_oracle_quadtree_structure(small_tree):
    ref, bmin, bmax = small_tree
    T_px = 32
    px = float(max(bmax[0] - bmin[0], bmax[1] - bmin[1])) / (4 * T_px) * 1.01  # two doublings below the box: deepest level 2
    info, tiles = ref.xray_quadtree(T_px, px, background=(255, 255, 255, 255))
    (rect, levels) = T.rect_and_levels(bmin, bmax, T_px, px)
    assert levels == 2 and info["deepest_level"] == 2 and (info["rect_min_x"], info["rect_min_y"], info["rect_edge"]) == rect
    ids = set(tiles)
    assert (0, 0) in ids and all((l - 1, i >> 2) in ids for (l, i) in ids if l > 0)  # every node's parent exists
    leaves = {k for k in ids if k[0] == 2}
    assert {(1, i >> 2) for _, i in leaves} == {k for k in ids if k[0] == 1}
    # a leaf equals the tile of its rectangle with the background assigned
    (l, i) = sorted(leaves)[0]
    r = T.quad_rect_of(l, i, rect)
    any_, img, _, _ = ref.xray_tile((r[0], r[1], bmin[2]), (r[0] + r[2], r[1] + r[2], bmax[2]), T_px, T_px)
    assert any_ and np.array_equal(tiles[(l, i)], O.assign_background(img, (255, 255, 255, 255)))
    # a parent equals build_parent + resize of its children
    p = (1, i >> 2)
    ch = [tiles.get((2, (p[1] << 2) + k)) for k in range(4)]
    assert np.array_equal(tiles[p], O.build_parent_tile(ch, (255, 255, 255, 255), T_px))
    # a sub-root: only its own subtree, rect of that node
    sub = sorted(k for k in ids if k[0] == 1)[0]
    info2, tiles2 = ref.xray_quadtree(T_px, px, root=sub)
    assert set(tiles2) == {k for k in ids if (k[0] == 1 and k == sub) or (k[0] == 2 and (1, k[1] >> 2) == sub)}
    assert (info2["rect_min_x"], info2["rect_min_y"], info2["rect_edge"]) == T.quad_rect_of(sub[0], sub[1], rect)
    assert all(np.array_equal(tiles2[k], tiles[k]) for k in tiles2)
    assert ref.xray_quadtree(T_px, px, root=(3, 0)) is None  # "Specified root node id is outside quadtree."


# ---- the quadtree's on-disk form: PNG tiles and the meta file (host code, csrc/xray_png.hpp) --------------------------------
def test_png_round_trip_through_pillow():
    import io

    Image = pytest.importorskip("PIL.Image")
    for img in (_noise(37, 1), _smooth(64, 2), np.full((5, 9, 4), 255, np.uint8), np.zeros((1, 1, 4), np.uint8)):
        raw = T.encode_png(img)
        assert raw[:8] == b"\x89PNG\r\n\x1a\n"
        back = Image.open(io.BytesIO(raw))
        back.load()
        assert back.mode == "RGBA" and back.size == (img.shape[1], img.shape[0])
        assert np.array_equal(np.asarray(back), img)


def test_xray_meta_against_python_protobuf():
    from proto_meta import XrayMeta

    nodes = [(0, 0), (1, 2), (2, 11), (3, int("301", 4))]
    raw = T.encode_xray_meta(300000.125, -200000.5, 1024.0, 3, 4096, nodes)
    m = XrayMeta.FromString(raw)
    assert m.version == 3 and m.deepest_level == 3 and m.tile_size == 4096  # CURRENT_VERSION, xray/src/lib.rs:15
    assert (m.bounding_rect.min.x, m.bounding_rect.min.y, m.bounding_rect.edge_length) == (300000.125, -200000.5, 1024.0)
    assert [(n.level, n.index) for n in m.nodes] == nodes and not m.bounding_rect.HasField("deprecated_min")
    assert m.SerializeToString() == raw
    assert [T.node_name(l, i) for l, i in nodes] == ["r", "r2", "r23", "r301"] == [pcv.xray_node_name(l, i) for l, i in nodes]


def test_fuzz_parent_tiles_and_binning():
    """hypothesis: arbitrary child / tile sizes (up- and down-scaling ratios), missing children, backgrounds; arbitrary columns
    for the binned aggregation - the product's shared code against the oracle / a dictionary."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(st.integers(1, 24), st.integers(1, 40), st.lists(st.booleans(), min_size=4, max_size=4), st.tuples(*[st.integers(0, 255)] * 4), st.integers(0, 2 ** 31))
    def parents(child_px, tile_px, present, bg, seed):
        hyp.assume(any(present))
        ch = [_noise(child_px, seed + k) if present[k] else None for k in range(4)]
        assert np.array_equal(T.build_parent(ch, bg, tile_px), O.build_parent_tile(ch, bg, tile_px))

    parents()

    vals = st.floats(-1e6, 1e6, allow_nan=False, width=32)

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(st.lists(st.tuples(st.integers(0, 15), vals, st.floats(0, 1, width=32)), min_size=1, max_size=200), st.sampled_from([0.5, 3.0, 1000.0, -7.0, 1e-3]))
    def bins(rows, size):
        pixel = np.array([r[0] for r in rows], np.uint32)
        attr = np.array([r[1] for r in rows], np.float32)
        value = np.array([[r[2]] for r in rows], np.float32)
        err, pix_sum, nb = T.binned(pixel, attr, value, size, 16, 1, bin_cap=256)
        want_sum, want_nb = _binned_reference(pixel, attr, value, size, 16)
        assert err == 0 and np.array_equal(nb, want_nb) and np.allclose(pix_sum, want_sum, rtol=1e-5, atol=1e-5)

    bins()
