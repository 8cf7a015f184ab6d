"""GPU parity of the X-ray pipeline beyond the leaf tile, through the C ABI against the oracle: binned columns,
assign_background, parent tiles (2 x 2 mosaic + Lanczos3 reduction) and the whole build_xray_quadtree
(xray/src/generation.rs:129-157, 410-451, 515-759).  (The file name sorts last on purpose: these entry points were added
after the last GPU session of their round.)"""
import numpy as np
import pytest

import oracle_api as O

pytestmark = pytest.mark.gpu

WHITE = (255, 255, 255, 255)
TRANSPARENT = (255, 255, 255, 0)


@pytest.fixture(scope="module")
def scene():
    import point_cloud_viewer_b200 as pcv

    n = 150_000
    x, y, z, rgb = pcv.synth_points_host(pcv.SYNTH_SLAB_ECEF, 80293751232, 0, n)
    inten = ((np.arange(n) * 7919) % 1000).astype(np.float32)
    bmin, bmax, res = pcv.synth_bbox(pcv.SYNTH_SLAB_ECEF)
    c = pcv.Context(0, max_points_per_node=4000)
    tree = c.build_octree(x, y, z, rgb, res, bmin, bmax, intensity=inten)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, intensity=inten, max_points_per_node=4000)
    yield dict(pcv=pcv, ctx=c, tree=tree, ref=ref, bmin=np.asarray(bmin), bmax=np.asarray(bmax))
    tree.free()
    c.close()


def _noise(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, n, 4), dtype=np.uint8)


def test_assign_background(scene):
    pcv, ctx = scene["pcv"], scene["ctx"]
    img = _noise(96, 1)
    img[::3, ::2, 3] = 127
    img[1::3, ::2, 3] = 128
    for bg in (WHITE, TRANSPARENT, (9, 8, 7, 6)):
        got = pcv.xray_assign_background(ctx, img.copy(), bg)
        assert np.array_equal(got, O.assign_background(img, bg))


@pytest.mark.parametrize("child_px,tile_px,missing", [(32, 32, ()), (32, 32, (0, 3)), (33, 33, (1,)), (16, 24, (2,)), (40, 16, (0, 1, 2)), (256, 256, ())])
def test_parent_tile(scene, child_px, tile_px, missing):
    """build_parent + image::imageops::resize(Lanczos3) (generation.rs:410-451, 722-759): every pixel equal to the oracle."""
    pcv, ctx = scene["pcv"], scene["ctx"]
    ch = [None if k in missing else _noise(child_px, 40 + k) for k in range(4)]
    for bg in (WHITE, TRANSPARENT):
        got = pcv.xray_build_parent(ctx, ch, bg, tile_px)
        assert np.array_equal(got, O.build_parent_tile(ch, bg, tile_px))


def test_binned_tiles(scene):
    """Binning = Some(("intensity", size)) for the colour and the intensity strategies.  The reference sums in arrival / hash
    order in f32, so: identical covered pixels, every channel within one grey level."""
    pcv, tree, ref = scene["pcv"], scene["tree"], scene["ref"]
    bmin, bmax = scene["bmin"], scene["bmax"]
    d = bmax - bmin
    tmin, tmax = bmin + [0.1, 0.1, 0.0] * d, bmin + [0.8, 0.8, 1.0] * d
    G = pcv.geometry
    q = G.quat_mul(G.quat_from_axis_angle([0, 0, 1], 0.7), G.quat_from_axis_angle([0, 1, 0], -0.9))
    qfg = G.Isometry((4157222.543, 664789.307, 4774952.099), q).inverse().as7()
    qmin, qmax = np.array([-40.0, -30.0, -10.0]), np.array([24.0, 34.0, 10.0])
    for mode, p0, p1 in ((pcv.XRAY_COLORED, 0.0, 0.0), (pcv.XRAY_INTENSITY, 0.0, 1000.0), (pcv.XRAY_INTENSITY, 50.0, 900.0)):
        for size in (100.0, 7.5, 1e9):
            for (lo, hi, w, h, frame) in ((tmin, tmax, 80, 56, None), (qmin, qmax, 96, 96, qfg)):
                any_g, got = tree.xray_tile_attr_binned(lo, hi, w, h, mode, size, p0, p1, query_from_global=frame)
                any_o, want = ref.xray_tile_attr_binned(lo, hi, w, h, mode, size, p0, p1, query_from_global=frame)
                assert any_g and any_o
                assert np.array_equal(got[..., 3], want[..., 3]), (mode, size, "covered pixels")
                diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
                assert diff.max() <= 1, (mode, size, int(diff.max()), int((diff > 0).sum()))
                assert (diff > 0).mean() < 0.03
                assert (got[..., 3] == 255).sum() > 100
        # one bin per column == Binning = None
        _, one = tree.xray_tile_attr_binned(tmin, tmax, 80, 56, mode, 1e9, p0, p1)
        _, none = tree.xray_tile_attr(tmin, tmax, 80, 56, mode, p0, p1, 0)
        assert np.abs(one.astype(np.int16) - none.astype(np.int16)).max() <= 1
    any_g, rgba = tree.xray_tile_attr_binned(bmax + 10, bmax + 20, 8, 8, pcv.XRAY_COLORED, 5.0)
    assert not any_g and (rgba == np.array(TRANSPARENT, np.uint8)).all()
    with pytest.raises(pcv._native.PcvError):
        tree.xray_tile_attr_binned(tmin, tmax, 8, 8, pcv.XRAY_HEIGHT_STDDEV, 5.0)


def test_quadtree_xray_strategy_equals_oracle(scene):
    """build_xray_quadtree with the XRay strategy: same node set, rect and levels, and every tile of every level
    byte-identical to the oracle's (leaf tiles are exact; parents follow from exact leaves)."""
    pcv, tree, ref = scene["pcv"], scene["tree"], scene["ref"]
    bmin, bmax = scene["bmin"], scene["bmax"]
    T = 64
    px = float(max(bmax[0] - bmin[0], bmax[1] - bmin[1])) / (4 * T) * 1.01  # deepest level 2: up to 16 leaves, 4 + 1 parents
    for bg in (WHITE, TRANSPARENT):
        info, tiles = tree.xray_quadtree(T, px, background=bg)
        oinfo, otiles = ref.xray_quadtree(T, px, background=bg)
        assert info["deepest_level"] == oinfo["deepest_level"] == 2
        assert (info["rect_min_x"], info["rect_min_y"], info["rect_edge"]) == (oinfo["rect_min_x"], oinfo["rect_min_y"], oinfo["rect_edge"])
        assert set(tiles) == set(otiles) and info["num_nodes"] == len(otiles)
        assert info["num_leaves"] == sum(1 for k in otiles if k[0] == 2) and info["kernel_launches"] > 0
        for k in sorted(otiles):
            assert np.array_equal(tiles[k], otiles[k]), k
    # with a query frame, from a sub-root, and cancellation from the callback
    G = pcv.geometry
    q = G.quat_mul(G.quat_from_axis_angle([0, 0, 1], 0.7), G.quat_from_axis_angle([0, 1, 0], -0.9))
    qfg = G.Isometry((4157222.543, 664789.307, 4774952.099), q).inverse().as7()
    info, tiles = tree.xray_quadtree(T, 1.0, query_from_global=qfg)
    oinfo, otiles = ref.xray_quadtree(T, 1.0, query_from_global=qfg)
    assert info["deepest_level"] == oinfo["deepest_level"] and set(tiles) == set(otiles)
    assert (info["rect_min_x"], info["rect_min_y"], info["rect_edge"]) == (oinfo["rect_min_x"], oinfo["rect_min_y"], oinfo["rect_edge"])
    for k in sorted(otiles):
        assert np.array_equal(tiles[k], otiles[k]), k
    sub = sorted(k for k in otiles if k[0] == 1)[0]
    info2, tiles2 = tree.xray_quadtree(T, 1.0, query_from_global=qfg, root=sub)
    oinfo2, otiles2 = ref.xray_quadtree(T, 1.0, query_from_global=qfg, root=sub)
    assert set(tiles2) == set(otiles2) and all(np.array_equal(tiles2[k], otiles2[k]) for k in otiles2)
    assert (info2["rect_min_x"], info2["rect_min_y"], info2["rect_edge"]) == (oinfo2["rect_min_x"], oinfo2["rect_min_y"], oinfo2["rect_edge"])
    seen = []
    with pytest.raises(pcv._native.PcvError) as e:
        tree.xray_quadtree(T, px, on_tile=lambda l, i, img: seen.append((l, i)) or len(seen) >= 2)
    assert e.value.code == -5 and len(seen) == 2
    with pytest.raises(pcv._native.PcvError):
        tree.xray_quadtree(T, px, root=(5, 0))  # "Specified root node id is outside quadtree."


def test_quadtree_other_strategies(scene):
    """The attribute strategies through the quadtree driver: leaves within one grey level of the oracle (f32 sums in
    unspecified order), same node set; parents are compared after rebuilding them from the product's own leaves."""
    pcv, tree, ref = scene["pcv"], scene["tree"], scene["ref"]
    bmin, bmax = scene["bmin"], scene["bmax"]
    T = 48
    px = float(max(bmax[0] - bmin[0], bmax[1] - bmin[1])) / (2 * T) * 1.01  # deepest level 1
    for kw in (dict(strategy=pcv.XRAY_COLORED), dict(strategy=pcv.XRAY_COLORED, bin_size=50.0), dict(strategy=pcv.XRAY_INTENSITY, p0=0.0, p1=1000.0, bin_size=20.0),
               dict(strategy=pcv.XRAY_HEIGHT_STDDEV, p0=1.5, colormap=1)):
        info, tiles = tree.xray_quadtree(T, px, background=TRANSPARENT, **kw)
        oinfo, otiles = ref.xray_quadtree(T, px, background=TRANSPARENT, **kw)
        assert info["deepest_level"] == 1 and set(tiles) == set(otiles)
        for k in (k for k in otiles if k[0] == 1):
            assert np.array_equal(tiles[k][..., 3], otiles[k][..., 3])
            assert np.abs(tiles[k].astype(np.int16) - otiles[k].astype(np.int16)).max() <= 1
        ch = [tiles.get((1, k)) for k in range(4)]
        assert np.array_equal(tiles[(0, 0)], O.build_parent_tile(ch, TRANSPARENT, T))
