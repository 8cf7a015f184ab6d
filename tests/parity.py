"""Shared parity helpers: compare a product octree (pcv.Octree or the test backend's tree) with the oracle's."""
import numpy as np

import oracle_api as O


def compare_trees(ref, tree, check_xyz=True):
    """ref: oracle_api.OracleOctree; tree: object with .nodes {name: dict(num_points, enc, cube, ...)} and
    .node_data(name) -> (xyz bytes, rgb bytes, intensity, src).  Bit-exact on node ids, counts, encodings,
    cubes, per-slot provenance (src index), colours, intensity and the stored position codes."""
    assert set(tree.nodes) == set(ref.nodes), "node sets differ: %s" % sorted(set(tree.nodes) ^ set(ref.nodes))[:8]
    for name, m in ref.nodes.items():
        t = tree.nodes[name]
        assert t["num_points"] == m["num_points"], (name, t["num_points"], m["num_points"])
        assert t["enc"] == m["enc"], (name, t["enc"], m["enc"])
        assert tuple(t["cube"]) == tuple(m["cube"]), (name, t["cube"], m["cube"])
        if m["num_points"] == 0:
            continue
        with_i = getattr(tree, "has_intensity", False)
        rx, rc, ri, rs = ref.node_data(name, with_i)
        tx, tc, ti, ts = tree.node_data(name)
        assert np.array_equal(rs, np.asarray(ts, np.uint64)), (name, "src index order")
        assert np.array_equal(rc.reshape(-1), np.asarray(tc).reshape(-1)), (name, "rgb")
        if with_i:
            assert np.array_equal(ri, ti), (name, "intensity")
        if check_xyz:
            assert np.array_equal(rx, tx), (name, "xyz codes")


def decode_node(meta, xyz_bytes):
    """Reference decode (codec.rs:124-139) in numpy for tolerance checks; returns (n,3) f64."""
    enc = meta["enc"]
    mn, edge = np.array(meta["cube"][:3]), meta["cube"][3]
    if enc == 1:
        v = np.frombuffer(xyz_bytes, np.uint8).reshape(-1, 3).astype(np.float64) / 255.0
    elif enc == 2:
        v = np.frombuffer(xyz_bytes, "<u2").reshape(-1, 3).astype(np.float64) / 65535.0
    elif enc == 3:
        v = np.frombuffer(xyz_bytes, "<f4").reshape(-1, 3).astype(np.float64)
    else:
        v = np.frombuffer(xyz_bytes, "<f8").reshape(-1, 3)
    return v * edge + mn
