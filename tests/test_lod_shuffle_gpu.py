"""the viewers draw every node in random order so that "the first N" points are a uniform subsample
(sdl_viewer/src/node_drawer.rs:185-205) and the reference asks for that order to be applied when the node is written
(src/octree/mod.rs:286-287).  pcv_octree_shuffle_nodes does it once on the GPU; every node must equal the oracle's
`reshuffle` (node_drawer.rs:34-43) of the unshuffled node under the product's keyed order, for positions and colours alike."""
import numpy as np
import pytest

import oracle_api as O

pytestmark = pytest.mark.gpu


def test_shuffled_nodes_equal_reference_reshuffle():
    import point_cloud_viewer_b200 as pcv

    n = 400_000
    x, y, z, rgb = O.synth_points(O.SYNTH_SLAB_ECEF, 9, 0, n)
    inten = (np.arange(n) % 977).astype(np.float32)
    bmin, bmax, res = O.synth_bbox(O.SYNTH_SLAB_ECEF)
    ctx = pcv.Context(0, max_points_per_node=3000)
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, intensity=inten, max_points_per_node=3000)
    tree = ctx.build_octree(x, y, z, rgb, res, bmin, bmax, intensity=inten)
    before = {nm: tree.node_data(nm) for nm, m in tree.nodes.items() if m["num_points"]}
    seed = 20240923
    tree.shuffle_nodes(seed)
    moved = 0
    encs = set()
    for nm, m in tree.nodes.items():
        cnt = m["num_points"]
        if not cnt:
            continue
        order = pcv.lod_order(seed, nm, cnt)
        assert np.array_equal(np.sort(order), np.arange(cnt, dtype=np.uint64)), nm  # a permutation
        bx, bc, bi, bs = before[nm]
        rx, rc, _, _ = ref.node_data(nm, True)
        assert np.array_equal(bx, rx) and np.array_equal(bc.reshape(-1), rc.reshape(-1))  # unshuffled == oracle (sanity)
        gx, gc, gi, gs = tree.node_data(nm)
        bpv = 3 * pcv.ENC_BYTES[m["enc"]]
        encs.add(m["enc"])
        assert np.array_equal(gx, O.reshuffle(order, rx, bpv)), (nm, "positions")
        assert np.array_equal(gc.reshape(-1), O.reshuffle(order, rc, 3)), (nm, "colours")
        assert np.array_equal(gs, bs[order.astype(np.int64)]) and np.array_equal(gi, bi[order.astype(np.int64)]), (nm, "provenance / intensity")
        moved += int((order != np.arange(cnt, dtype=np.uint64)).sum())
    assert moved > 0.9 * n and len(encs) >= 2
    # different nodes get different orders; the same seed reproduces the order
    big = [nm for nm, m in tree.nodes.items() if m["num_points"] >= 1000][:2]
    a, b = pcv.lod_order(seed, big[0], 1000), pcv.lod_order(seed, big[1], 1000)
    assert not np.array_equal(a, b) and np.array_equal(a, pcv.lod_order(seed, big[0], 1000))
    # queries still work on the shuffled octree: same survivor set as before the shuffle
    G = pcv.geometry
    loc = G.aabb(bmin + 0.2 * (bmax - bmin), bmin + 0.7 * (bmax - bmin))
    got = np.sort(np.concatenate([bt["src"] for bt in tree.query_points(loc)] or [np.zeros(0, np.uint64)]))
    ol = O.Location()
    for f, _ in O.Location._fields_:
        setattr(ol, f, getattr(loc, f))
    assert np.array_equal(got, np.sort(ref.query(ol)["src"]))
    tree.free()
    ctx.close()
