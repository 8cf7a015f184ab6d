"""`/nodes_data` reply blob (octree_web_viewer/src/backend.rs:66-75, 92-165).
CPU: the oracle's blob has the documented layout (the JavaScript client's parsing rules).  GPU: the product's blob,
gathered on the device, is byte-identical to the oracle's."""
import struct

import numpy as np
import pytest

import oracle_api as O

ENC_BYTES = {1: 1, 2: 2, 3: 4, 4: 8}


def _points(n, seed):
    rng = np.random.default_rng(seed)
    cen = rng.random((9, 3)) * [120, 120, 12]
    P = cen[rng.integers(0, 9, n)] + rng.normal(0, 1.2, (n, 3)) + [4.1e6, 6.6e5, 4.7e6]
    rgb = rng.integers(0, 255, (n, 3), dtype=np.uint8)
    return P, rgb


def _parse(blob, count):
    """What the web client does: header (min xyz, edge, n u32, bpc u8), pad 8, positions, pad 8, colours, pad 8."""
    out, o = [], 0
    for _ in range(count):
        mx, my, mz, edge, n, bpc = struct.unpack_from("<ddddIB", blob, o)
        o += 40
        pos = blob[o : o + n * 3 * bpc]
        o += (n * 3 * bpc + 7) // 8 * 8
        col = blob[o : o + n * 3]
        o += (n * 3 + 7) // 8 * 8
        out.append(((mx, my, mz), edge, n, bpc, pos, col))
    assert o == len(blob)
    return out


def test_oracle_blob_layout():
    P, rgb = _points(60000, 1)
    ref = O.build(*[np.ascontiguousarray(P[:, i]) for i in range(3)], rgb, 1e-4, P.min(0), P.max(0), max_points_per_node=900)
    names = [nm for nm, m in ref.nodes.items() if m["num_points"] > 0][:40]
    blob = ref.nodes_data_blob(names)
    assert len(blob) % 8 == 0
    for nm, (mn, edge, n, bpc, pos, col) in zip(names, _parse(blob, len(names))):
        m = ref.nodes[nm]
        assert n == m["num_points"] and bpc == ENC_BYTES[m["enc"]] and mn == tuple(m["cube"][:3]) and edge == m["cube"][3]
        xyz, c, _, _ = ref.node_data(nm)
        assert pos == xyz.tobytes() and col == c.tobytes()
    empty = [nm for nm, m in ref.nodes.items() if m["num_points"] == 0]
    with pytest.raises(KeyError):
        ref.nodes_data_blob(["r7777777"])  # unknown id: NodeNotFound
    if empty:
        with pytest.raises(KeyError):
            ref.nodes_data_blob([names[0], empty[0]])  # files of zero-point nodes are deleted: NodeNotFound
    assert ref.nodes_data_blob([]) == b""


@pytest.mark.gpu
@pytest.mark.parametrize("n,maxpts,res", [(200000, 700, 1e-4), (120000, 5000, 1e-9), (30000, 100000, 1e-3)])
def test_blob_matches_oracle(n, maxpts, res):
    import point_cloud_viewer_b200 as pcv

    P, rgb = _points(n, n)
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    c = pcv.Context(0, max_points_per_node=maxpts)
    tree = c.build_octree(x, y, z, rgb.reshape(-1), res, P.min(0), P.max(0))
    ref = O.build(x, y, z, rgb, res, P.min(0), P.max(0), max_points_per_node=maxpts)
    full = [nm for nm, m in tree.nodes.items() if m["num_points"] > 0]
    rng = np.random.default_rng(0)
    for names in (full, full[::-1][:17], [full[i] for i in rng.permutation(len(full))[:5]], full[:1], []):
        assert tree.nodes_data_blob(names).tobytes() == ref.nodes_data_blob(names)
    # visible nodes of a camera, as the viewer asks for them
    G = pcv.geometry
    M = G.Perspective.new_fov(1.0, 1.2, 0.1, 500.0).matrix @ G.Isometry((P.mean(0) + [0, 0, 80]).tolist(), (0, 0, 0, 1)).inverse().to_homogeneous()
    vis = tree.get_visible_nodes(M)
    assert vis == ref.visible_nodes(M.T.reshape(-1)) and len(vis) > 0
    assert tree.nodes_data_blob(vis).tobytes() == ref.nodes_data_blob(vis)
    with pytest.raises(pcv.PcvError) as e:
        tree.nodes_data_blob(["r7777777"])
    assert e.value.code == -4
    empty = [nm for nm, m in tree.nodes.items() if m["num_points"] == 0]
    if empty:
        with pytest.raises(pcv.PcvError) as e:
            tree.nodes_data_blob([full[0], empty[0]])
        assert e.value.code == -4
    tree.free()
    c.close()
