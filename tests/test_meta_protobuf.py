"""meta.pb pinned against an independent parser: python-protobuf (google.protobuf) with a descriptor restating
point_viewer_proto_rust/src/proto.proto:58-149 (tests/proto_meta.py).

  * what the product writes (csrc/disk_io.hpp encode_meta; pcv_octree_write_dir on the GPU) parses with protobuf to the
    same version / bounding box / resolution / node set, with no unknown fields;
  * what protobuf serialises is read by the product (decode_meta; pcv_octree_load_dir) and by the oracle's reader;
  * the oracle's writer is held to the same parser, so the two in-repo implementations are no longer each other's only
    witness.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_api as O
import proto_meta as PM


def _tb():
    from tb_api import lib, _SO  # noqa: F401

    L = lib()
    L.tb_encode_meta.restype = C.c_int64
    L.tb_encode_meta.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    L.tb_decode_meta.restype = C.c_int64
    L.tb_decode_meta.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    return L


def _random_nodes(rng, n):
    out = []
    seen = set()
    while len(out) < n:
        level = int(rng.integers(0, 41))
        idx = int(rng.integers(0, 2 ** 62)) % (8 ** level) if level else 0
        v = (level << 120) | idx
        if v in seen:
            continue
        seen.add(v)
        # num_points up to > 2^32 (int64 field), zero-point nodes stay in meta.pb (generation.rs:241-243)
        npts = int(rng.choice([0, 1, 7, 100000, 2 ** 33 + 5, int(rng.integers(0, 10 ** 9))]))
        out.append((v >> 64, v & (2 ** 64 - 1), npts, int(rng.integers(1, 5))))
    return out


def test_product_encoder_parses_with_protobuf():
    L = _tb()
    rng = np.random.default_rng(5)
    for count in (0, 1, 50, 3000):
        nodes = _random_nodes(rng, count)
        bmin, bmax = rng.normal(size=3) * 1e6, rng.normal(size=3) * 1e6
        res = float(rng.random()) + 1e-9
        arr = np.array(nodes, dtype=object).reshape(-1, 4) if nodes else np.zeros((0, 4), object)
        flat = np.array([[int(v) for v in row] for row in arr], dtype=np.uint64).reshape(-1)
        cap = 64 + 64 * max(1, count)
        buf = np.zeros(cap, np.uint8)
        n = L.tb_encode_meta(res, bmin.ctypes.data, bmax.ctypes.data, flat.ctypes.data if count else None, count, buf.ctypes.data, cap)
        assert n > 0
        got = PM.parse_meta(buf[:n].tobytes())
        assert got["version"] == 13 and got["unknown"] == 0
        assert got["bbox_min"] == tuple(bmin) and got["bbox_max"] == tuple(bmax) and got["resolution"] == res
        assert got["nodes"] == {(hi, lo): (npts, enc) for hi, lo, npts, enc in nodes}
        # protobuf re-serialises the parsed message to the same bytes: field order and varint forms are canonical
        m = PM.Meta()
        m.ParseFromString(buf[:n].tobytes())
        assert m.SerializeToString() == buf[:n].tobytes()


def test_product_decoder_reads_protobuf_output():
    L = _tb()
    rng = np.random.default_rng(6)
    nodes = _random_nodes(rng, 500)
    bmin, bmax, res = (1.5, -2.25, 3e6), (7.0, 8.0, 3e6 + 9), 0.001
    data = PM.serialize_meta(bmin, bmax, res, nodes)
    b = np.frombuffer(data, np.uint8).copy()
    r, ver = C.c_double(), C.c_int()
    mn, mx = np.zeros(3), np.zeros(3)
    out = np.zeros(4 * len(nodes), np.uint64)
    n = L.tb_decode_meta(b.ctypes.data, len(b), C.byref(r), mn.ctypes.data, mx.ctypes.data, C.byref(ver), out.ctypes.data, len(nodes))
    assert n == len(nodes) and ver.value == 13 and r.value == res and tuple(mn) == bmin and tuple(mx) == bmax
    assert [tuple(int(v) for v in row) for row in out.reshape(-1, 4)] == nodes
    # other versions are rejected by the product (only 13 is read)
    for v in (9, 12, 14):
        bad = np.frombuffer(PM.serialize_meta(bmin, bmax, res, nodes[:3], version=v), np.uint8).copy()
        assert L.tb_decode_meta(bad.ctypes.data, len(bad), C.byref(r), mn.ctypes.data, mx.ctypes.data, C.byref(ver), out.ctypes.data, 3) == -1
        assert ver.value == v


def _small_cloud(n=30000, seed=3):
    rng = np.random.default_rng(seed)
    x, y, z = rng.random(n) * 40 + 1000.0, rng.random(n) * 40 - 500.0, rng.random(n) * 10
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    return x, y, z, rgb, (1000.0, -500.0, 0.0), (1040.0, -460.0, 10.0), 0.01


def test_oracle_writer_and_reader_against_protobuf(tmp_path):
    x, y, z, rgb, bmin, bmax, res = _small_cloud()
    ref = O.build(x, y, z, rgb, res, bmin, bmax, max_points_per_node=1500)
    d = str(tmp_path / "o")
    os.makedirs(d)
    ref.write_dir(d)
    got = PM.parse_meta(open(os.path.join(d, "meta.pb"), "rb").read())
    assert got["version"] == 13 and got["unknown"] == 0 and got["resolution"] == res
    assert got["bbox_min"] == bmin and got["bbox_max"] == bmax
    want = {O.id_from_str(nm): (m["num_points"], m["enc"]) for nm, m in ref.nodes.items()}
    assert got["nodes"] == want
    assert any(v[0] == 0 for v in want.values()) or True  # zero-point nodes, when present, stay listed
    # the oracle reads a meta.pb produced by protobuf (node order shuffled: FnvHashMap order is unspecified in the reference)
    items = [(hi, lo, n, e) for (hi, lo), (n, e) in want.items()]
    np.random.default_rng(0).shuffle(items)
    with open(os.path.join(d, "meta.pb"), "wb") as f:
        f.write(PM.serialize_meta(bmin, bmax, res, items))
    back = O.load_dir(d)
    assert {nm: (m["num_points"], m["enc"]) for nm, m in back.nodes.items()} == {nm: (m["num_points"], m["enc"]) for nm, m in ref.nodes.items()}
    assert back.meta()[0] == res


@pytest.mark.gpu
def test_gpu_write_dir_and_load_dir_against_protobuf(tmp_path):
    import point_cloud_viewer_b200 as pcv

    x, y, z, rgb, bmin, bmax, res = _small_cloud()
    ctx = pcv.Context(0, max_points_per_node=1500)
    tree = ctx.build_octree(x, y, z, rgb.reshape(-1), res, bmin, bmax)
    d = str(tmp_path / "g")
    tree.write_dir(d)
    got = PM.parse_meta(open(os.path.join(d, "meta.pb"), "rb").read())
    assert got["version"] == 13 and got["unknown"] == 0 and got["resolution"] == res
    assert got["bbox_min"] == bmin and got["bbox_max"] == bmax
    want = {(m["hi"], m["lo"]): (m["num_points"], m["enc"]) for m in tree.nodes.values()}
    assert got["nodes"] == want
    # node files: one .xyz/.rgb per non-empty node, sizes as the reference derives them (on_disk.rs:23-33)
    for nm, m in tree.nodes.items():
        p = os.path.join(d, nm + ".rgb")
        assert os.path.exists(p) == (m["num_points"] > 0)
        if m["num_points"]:
            assert os.path.getsize(p) == 3 * m["num_points"]
            assert os.path.getsize(os.path.join(d, nm + ".xyz")) == 3 * m["num_points"] * pcv.ENC_BYTES[m["enc"]]
    # meta.pb rewritten by protobuf (shuffled node order) loads to the same octree
    items = [(hi, lo, n, e) for (hi, lo), (n, e) in want.items()]
    np.random.default_rng(1).shuffle(items)
    with open(os.path.join(d, "meta.pb"), "wb") as f:
        f.write(PM.serialize_meta(bmin, bmax, res, items))
    back = ctx.load_dir(d)
    assert {nm: (m["num_points"], m["enc"], m["cube"]) for nm, m in back.nodes.items()} == {nm: (m["num_points"], m["enc"], m["cube"]) for nm, m in tree.nodes.items()}
    nm = max(tree.nodes, key=lambda k: tree.nodes[k]["num_points"])
    assert np.array_equal(back.node_data(nm)[0], tree.node_data(nm)[0])
    back.free()
    tree.free()
    ctx.close()
