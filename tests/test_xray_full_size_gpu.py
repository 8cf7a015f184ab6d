"""BASELINE config 5 at its stated tile size (missing #1): 4096 x 4096 X-ray leaf tiles over an octree of the
benchmark generator, compared with the oracle's xray_from_points restatement (xray/src/generation.rs:464-513) - the per-pixel
z-bucket sets and the RGBA image, plain and with a query_from_global isometry (OBB location + transformed points, :471-497)."""
import numpy as np
import pytest

import oracle_api as O

pytestmark = pytest.mark.gpu


def test_4096_tiles_equal_oracle():
    import point_cloud_viewer_b200 as pcv

    n = 12_000_000
    kind = pcv.SYNTH_GAUSS_CLUSTERS
    x, y, z, rgb = O.synth_points(kind, 1, 0, n)
    bmin, bmax, res = O.synth_bbox(kind)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax)
    ctx = pcv.Context(0)
    tree = ctx.build_octree(x, y, z, rgb, res, bmin, bmax)
    G = pcv.geometry
    W = 4096
    # (1) the whole cube in one tile (every node, 0.25 m pixels), (2) one of bench.py's 16 leaf tiles (256 m, 0.0625 m pixels)
    for tmin, tmax in ((bmin, bmax), (bmin + [256.0, 512.0, 0.0], bmin + [512.0, 768.0, 1024.0])):
        any_g, rgba, zb = tree.xray_tile(tmin, tmax, W, W, want_bits=True)
        any_o, rgba_o, zb_o, _ = ref.xray_tile(tmin, tmax, W, W)
        assert any_g and any_o
        assert np.array_equal(rgba, rgba_o)
        assert np.array_equal(zb, zb_o)
        assert (rgba[..., 3] == 255).sum() > 10000
        st = tree.last_xray_stats()
        assert st["points"] > 0 and st["algorithmic_bytes"] >= 4 * W * W
        del zb, zb_o
    # (3) a rotated query frame around the cube centre: OBB culling + per-point transform
    centre = (bmin + bmax) * 0.5
    q = G.quat_mul(G.quat_from_axis_angle([0, 0, 1], 0.4), G.quat_from_axis_angle([1, 0, 0], 0.15))
    query_from_global = G.Isometry(centre, q).inverse().as7()
    qmin, qmax = np.array([-300.0, -300.0, -520.0]), np.array([300.0, 300.0, 520.0])
    any_g, rgba, zb = tree.xray_tile(qmin, qmax, W, W, query_from_global=query_from_global, want_bits=True)
    any_o, rgba_o, zb_o, _ = ref.xray_tile(qmin, qmax, W, W, query_from_global=query_from_global)
    assert any_g and any_o and np.array_equal(rgba, rgba_o) and np.array_equal(zb, zb_o)
    assert (rgba[..., 3] == 255).sum() > 10000
    tree.free()
    ctx.close()
