"""PLY input path, CPU side: the oracle restatement (oracle/oracle_ply.hpp) against the reference's own fixtures and
unit-test assertions (src/read_write/ply.rs:746-790, committed as tests/golden/ply_fixtures.json), and the product's
header parser (pcv_ply_read_header: host code, no GPU involved) against the oracle on generated layouts."""
import json
import os

import numpy as np
import pytest

import oracle_api as O
from ply_util import write_ply

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, "golden", "ply_fixtures.json")))


@pytest.fixture(params=sorted(FIX))
def fixture_file(request, tmp_path):
    p = tmp_path / request.param
    p.write_bytes(bytes.fromhex(FIX[request.param]["hex"]))
    return str(p), FIX[request.param]["expect"]


def test_oracle_reads_the_reference_fixtures(fixture_file):
    path, exp = fixture_file
    info = O.ply_open(path)
    assert info["num_points"] == exp["num_points"] and info["has_color"] and info["has_intensity"] == exp["intensity"]
    batches = O.ply_batches(path, 2)  # BATCH_SIZE = 2 (ply.rs:738-740)
    assert len(batches) == exp["batches_of_2"]
    assert batches[0][0][0] == exp["first_x"] and batches[-1][0][-1] == exp["last_x"]
    assert batches[0][3][0, 0] == exp["first_red"] and batches[-1][3][-1, 0] == exp["last_red"]
    if exp["intensity"]:
        assert len(batches[0][4]) == 2 and len(batches[-1][4]) == 2
        assert all(np.isnan(b[4]).all() for b in batches)
    # batch boundaries do not change values
    whole = O.ply_read(path)
    assert np.array_equal(np.concatenate([b[0] for b in batches]), whole[0]) and np.array_equal(np.concatenate([b[3] for b in batches]), whole[3])


def test_product_header_matches_oracle_on_fixtures(fixture_file):
    import point_cloud_viewer_b200 as pcv

    path, exp = fixture_file
    info = pcv.ply_read_header(path)
    o = O.ply_open(path)
    assert info.num_points == o["num_points"] == exp["num_points"]
    assert info.header_bytes == o["header_bytes"] and info.record_bytes == o["record_bytes"]
    assert bool(info.has_color) == o["has_color"] and bool(info.has_intensity) == o["has_intensity"]


LAYOUTS = [
    [("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "red"), ("uchar", "green"), ("uchar", "blue")],
    [("double", "x"), ("double", "y"), ("double", "z"), ("uchar", "r"), ("uchar", "g"), ("uchar", "b"), ("uchar", "alpha"), ("float", "intensity")],
    [("uchar", "red"), ("short", "junk"), ("float", "z"), ("uchar", "green"), ("int", "y"), ("ushort", "x"), ("uchar", "blue"), ("uint", "skipme"), ("float", "intensity"),
     ("double", "weight"), ("uchar", "classification"), ("char", "tiny")],
    [("char", "x"), ("uchar", "y"), ("short", "z"), ("uchar", "red"), ("uchar", "green"), ("uchar", "blue"), ("float", "a")],
    [("uint", "x"), ("int", "y"), ("double", "z")],
]


@pytest.mark.parametrize("li", range(len(LAYOUTS)))
def test_product_header_matches_oracle_on_layouts(li, tmp_path):
    import point_cloud_viewer_b200 as pcv

    rng = np.random.default_rng(li)
    path = str(tmp_path / "a.ply")
    off = (1.5e6, -2.25e5, 12.125) if li % 2 else None
    write_ply(path, 17, LAYOUTS[li], rng, offset=off, comments=("made by a test", "offset: 1 2"))
    info, o = pcv.ply_read_header(path), O.ply_open(path)
    assert (info.num_points, info.header_bytes, info.record_bytes) == (o["num_points"], o["header_bytes"], o["record_bytes"])
    assert bool(info.has_color) == o["has_color"] and bool(info.has_intensity) == o["has_intensity"]
    assert tuple(info.offset) == o["offset"] == (tuple(float(v) for v in off) if off else (0.0, 0.0, 0.0))
    # the oracle's field table: role 1,2,3 = x,y,z; 4,5,6 = r,g,b; 7 = intensity
    fields = {role: (typ, offset) for role, typ, offset, _ in O.ply_fields(path) if role in (1, 2, 3, 4, 5, 6, 7)}
    for a in range(3):
        assert (info.type_xyz[a], info.off_xyz[a]) == fields[1 + a]
        if info.has_color:
            assert info.off_rgb[a] == fields[4 + a][1]
    if info.has_intensity:
        assert info.off_intensity == fields[7][1]


def test_oracle_semantics_on_generated_file(tmp_path):
    """as-f64 casts (int8 read unsigned, ply.rs:254), offset added, alpha skipped as one byte, skipped property sizes."""
    rng = np.random.default_rng(5)
    path = str(tmp_path / "b.ply")
    cols = write_ply(path, 1000, LAYOUTS[3], rng, offset=(10.0, 20.0, 30.0))
    # "a" is declared float but skipped as ONE byte (ply.rs:383-385): the reader's record is 3 bytes shorter than the
    # writer's, so (as in the reference) only the first record lines up
    assert O.ply_open(path)["record_bytes"] == 1 + 1 + 2 + 3 + 1
    x, y, z, rgb, inten = O.ply_read(path, 0, 1)
    assert x[0] == float(cols["x"].view(np.uint8)[0]) + 10.0  # sic: int8 is read as an unsigned byte (ply.rs:254)
    assert y[0] == float(cols["y"][0]) + 20.0 and z[0] == float(cols["z"][0]) + 30.0
    assert inten is None and rgb[0, 1] == cols["green"][0]
    # the same layout with a one-byte alpha reads back completely
    props = LAYOUTS[3][:-1] + [("uchar", "a")]
    cols = write_ply(path, 1000, props, rng, offset=(10.0, 20.0, 30.0))
    x, y, z, rgb, inten = O.ply_read(path)
    assert np.array_equal(x, cols["x"].view(np.uint8).astype(np.float64) + 10.0)
    assert np.array_equal(y, cols["y"].astype(np.float64) + 20.0) and np.array_equal(z, cols["z"].astype(np.float64) + 30.0)
    assert np.array_equal(rgb[:, 1], cols["green"])
    mn, mx = O.ply_find_bounding_box(path)
    assert mn == (x.min(), y.min(), z.min()) and mx == (x.max(), y.max(), z.max())


def test_bounding_box_and_empty_file(tmp_path):
    rng = np.random.default_rng(6)
    path = str(tmp_path / "c.ply")
    cols = write_ply(path, 5000, LAYOUTS[1], rng, offset=(4.0e6, 5.0e5, 4.5e6))
    x, y, z, rgb, inten = O.ply_read(path)
    assert np.array_equal(x, cols["x"] + 4.0e6) and np.array_equal(inten, cols["intensity"])
    mn, mx = O.ply_find_bounding_box(path)
    assert mn == (x.min(), y.min(), z.min()) and mx == (x.max(), y.max(), z.max())
    empty = str(tmp_path / "e.ply")
    write_ply(empty, 0, LAYOUTS[0], rng)
    assert O.ply_find_bounding_box(empty) == ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))  # Aabb::zero, generation.rs:269


BAD_HEADERS = [
    (b"plx\nformat binary_little_endian 1.0\nend_header\n", "Not a PLY"),
    (b"ply\nformat binary_little_endian 2.0\nend_header\n", "Invalid version"),
    (b"ply\nformat weird 1.0\nend_header\n", "Invalid format"),
    (b"ply\nformat binary_little_endian 1.0\nproperty float x\nend_header\n", "outside of element"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex abc\nend_header\n", "Invalid count"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty quad x\nend_header\n", "Invalid data type"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\n\nend_header\n", "Invalid line"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nobj_info hi\nend_header\n", "Invalid line"),
    (b"ply\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n", "No format"),
    (b"ply\nformat binary_little_endian 1.0\nelement face 1\nend_header\n", "vertex"),
    (b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n", "nsupported"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\nproperty float y\nend_header\n", "'x', 'y', 'z'"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nproperty float normal0\nend_header\n", "Multidimensional"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\n", "Invalid line"),
    (b"ply\nformat binary_little_endian 1.0\ncomment offset: 1 2 x\nelement vertex 1\nend_header\n", "Invalid offset"),
]


@pytest.mark.parametrize("bi", range(len(BAD_HEADERS)))
def test_header_errors_match(bi, tmp_path):
    """Every error condition of parse_header / from_file: the oracle and the product reject the same files."""
    import point_cloud_viewer_b200 as pcv

    data, needle = BAD_HEADERS[bi]
    path = str(tmp_path / "bad.ply")
    open(path, "wb").write(data)
    with pytest.raises(O.PlyError) as eo:
        O.ply_open(path)
    assert needle in str(eo.value)
    with pytest.raises(pcv.PcvError) as ep:
        pcv.ply_read_header(path)
    assert needle in str(ep.value)


def test_missing_file():
    import point_cloud_viewer_b200 as pcv

    with pytest.raises(O.PlyError):
        O.ply_open("/nonexistent/x.ply")
    with pytest.raises(pcv.PcvError) as e:
        pcv.ply_read_header("/nonexistent/x.ply")
    assert e.value.code == -3
