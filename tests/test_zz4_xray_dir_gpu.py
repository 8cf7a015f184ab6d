"""build_xray_quadtree with the reference's outputs (xray/src/generation.rs:560-622): <dir>/<node id>.png read back with Pillow
equal the tiles the callback form delivers (and therefore the oracle's), the meta file parsed by python-protobuf.
(Sorts last: added after the round's last GPU session.)"""
import os

import numpy as np
import pytest

import oracle_api as O

pytestmark = pytest.mark.gpu


def test_quadtree_directory(ctx, tmp_path):
    import point_cloud_viewer_b200 as pcv
    from proto_meta import XrayMeta

    Image = pytest.importorskip("PIL.Image")
    n = 120_000
    x, y, z, rgb = pcv.synth_points_host(pcv.SYNTH_SLAB_ECEF, 80293751232, 0, n)
    bmin, bmax, res = pcv.synth_bbox(pcv.SYNTH_SLAB_ECEF)
    tree = ctx.build_octree(x, y, z, rgb, res, bmin, bmax)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax)
    T = 64
    px = float(max(bmax[0] - bmin[0], bmax[1] - bmin[1])) / (4 * T) * 1.01
    for root, meta_name in (((0, 0), "meta.pb"), (None, None)):
        if root is None:  # a sub-root: the first level-1 node that exists
            root = sorted(k for k in otiles if k[0] == 1)[0]
            meta_name = "meta%d.pb" % root[1]
        d = str(tmp_path / ("q%d_%d" % root))
        info = tree.xray_quadtree_write_dir(d, T, px, root=root)
        oinfo, otiles_r = ref.xray_quadtree(T, px, root=root)
        if root == (0, 0):
            otiles = otiles_r
        names = {pcv.xray_node_name(l, i): (l, i) for (l, i) in otiles_r}
        assert set(os.listdir(d)) == {nm + ".png" for nm in names} | {meta_name}
        for nm, key in names.items():
            img = Image.open(os.path.join(d, nm + ".png"))
            assert img.mode == "RGBA" and np.array_equal(np.asarray(img), otiles_r[key]), nm
        m = XrayMeta.FromString(open(os.path.join(d, meta_name), "rb").read())
        assert m.version == 3 and m.tile_size == T and m.deepest_level == oinfo["deepest_level"] == info["deepest_level"]
        assert (m.bounding_rect.min.x, m.bounding_rect.min.y, m.bounding_rect.edge_length) == (oinfo["rect_min_x"], oinfo["rect_min_y"], oinfo["rect_edge"])
        assert {(k.level, k.index) for k in m.nodes} == set(otiles_r) and len(m.nodes) == info["num_nodes"]
    tree.free()


def test_full_size_parent_tile(ctx):
    """A 2048 x 2048 parent from four 2048 x 2048 children (one missing): more blocks than one wave of the grid-stride kernels,
    byte-identical to the oracle."""
    import point_cloud_viewer_b200 as pcv

    rng = np.random.default_rng(5)
    T = 2048
    ch = [rng.integers(0, 256, (T, T, 4), dtype=np.uint8) if k != 2 else None for k in range(4)]
    got = pcv.xray_build_parent(ctx, ch, (255, 255, 255, 0), T)
    assert np.array_equal(got, O.build_parent_tile(ch, (255, 255, 255, 0), T))
    img = rng.integers(0, 256, (3000, 3000, 4), dtype=np.uint8)
    assert np.array_equal(pcv.xray_assign_background(ctx, img.copy(), (1, 2, 3, 4)), O.assign_background(img, (1, 2, 3, 4)))
