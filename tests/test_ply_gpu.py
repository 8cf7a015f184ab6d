"""GPU parity of the PLY input path (through the C ABI) against the oracle restatement of PlyIterator /
find_bounding_box / build_octree_from_file: bit-exact positions, colours, intensities, bounding box and octree."""
import json
import os

import numpy as np
import pytest

import oracle_api as O
from parity import compare_trees
from ply_util import write_ply
from test_ply_pins import FIX, LAYOUTS

pytestmark = pytest.mark.gpu


def _device_points(pp):
    x, y, z = (b.tensor()[: pp.n].cpu().numpy() for b in (pp.x, pp.y, pp.z))
    rgb = pp.rgb.tensor()[: 3 * pp.n].cpu().numpy().reshape(-1, 3) if pp.rgb else None
    inten = pp.intensity.tensor()[: pp.n].cpu().numpy() if pp.intensity else None
    return x, y, z, rgb, inten


def _same(a, b):
    return (a is None and b is None) or np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", sorted(FIX))
def test_reference_fixtures(ctx, name, tmp_path):
    """src/read_write/ply.rs:746-790 through the GPU path."""
    path = str(tmp_path / name)
    open(path, "wb").write(bytes.fromhex(FIX[name]["hex"]))
    exp = FIX[name]["expect"]
    pp = ctx.load_ply(path)
    batches = list(pp.batches(2))
    assert len(batches) == exp["batches_of_2"]
    assert batches[0]["position"][0, 0] == exp["first_x"] and batches[-1]["position"][-1, 0] == exp["last_x"]
    assert batches[0]["color"][0, 0] == exp["first_red"] and batches[-1]["color"][-1, 0] == exp["last_red"]
    if exp["intensity"]:
        assert len(batches[0]["intensity"]) == 2 and all(np.isnan(b["intensity"]).all() for b in batches)
    got, ref = _device_points(pp), O.ply_read(path)
    assert all(_same(g, r) for g, r in zip(got, ref))
    assert (tuple(pp.bbox_min), tuple(pp.bbox_max)) == O.ply_find_bounding_box(path)
    pp.free()


@pytest.mark.parametrize("li,n", [(0, 100000), (1, 70001), (2, 33333), (3, 5000), (4, 4097), (0, 1), (1, 4096), (2, 255)])
def test_layouts_match_oracle(ctx, li, n, tmp_path):
    rng = np.random.default_rng(100 * li + n)
    props = LAYOUTS[li] if li != 3 else LAYOUTS[3][:-1] + [("uchar", "a")]
    path = str(tmp_path / "l.ply")
    write_ply(path, n, props, rng, offset=(4.1e6, 6.6e5, -4.7e6) if li % 2 == 0 else None)
    pp = ctx.load_ply(path)
    got, ref = _device_points(pp), O.ply_read(path)
    for g, r in zip(got, ref):
        assert _same(g, r)
    assert (tuple(pp.bbox_min), tuple(pp.bbox_max)) == O.ply_find_bounding_box(path)
    pp.free()


def test_multi_chunk_file(ctx, tmp_path):
    """> 64 MiB of records: several trips through the pinned ring; odd record size (15 bytes)."""
    n = 6_000_011
    rng = np.random.default_rng(1)
    path = str(tmp_path / "big.ply")
    cols = write_ply(path, n, LAYOUTS[0], rng, offset=(1.0e6, 2.0e6, 3.0e6))
    pp = ctx.load_ply(path)
    x, y, z, rgb, _ = _device_points(pp)
    assert np.array_equal(x, cols["x"].astype(np.float64) + 1.0e6) and np.array_equal(y, cols["y"].astype(np.float64) + 2.0e6)
    assert np.array_equal(z, cols["z"].astype(np.float64) + 3.0e6)
    assert np.array_equal(rgb, np.stack([cols["red"], cols["green"], cols["blue"]], 1))
    assert tuple(pp.bbox_min) == (x.min(), y.min(), z.min()) and tuple(pp.bbox_max) == (x.max(), y.max(), z.max())
    pp.free()


def test_unpack_kernel_entry(ctx, tmp_path):
    """pcv_ply_unpack_device on records already resident in device memory."""
    import torch

    import point_cloud_viewer_b200 as pcv

    n = 50000
    rng = np.random.default_rng(2)
    path = str(tmp_path / "u.ply")
    write_ply(path, n, LAYOUTS[2], rng)
    info = pcv.ply_read_header(path)
    body = np.fromfile(path, np.uint8, offset=info.header_bytes, count=n * info.record_bytes)
    raw = torch.from_numpy(body).cuda()
    x, y, z = (torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3))
    rgb = torch.empty(3 * n, dtype=torch.uint8, device="cuda")
    inten = torch.empty(n, dtype=torch.float32, device="cuda")
    mn, mx = ctx.ply_unpack_device(info, raw.data_ptr(), n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), inten.data_ptr())
    torch.cuda.synchronize()
    rx, ry, rz, rrgb, rint = O.ply_read(path)
    assert np.array_equal(x.cpu().numpy(), rx) and np.array_equal(y.cpu().numpy(), ry) and np.array_equal(z.cpu().numpy(), rz)
    assert np.array_equal(rgb.cpu().numpy().reshape(-1, 3), rrgb) and np.array_equal(inten.cpu().numpy(), rint)
    assert (tuple(mn), tuple(mx)) == O.ply_find_bounding_box(path)


@pytest.mark.parametrize("with_intensity", [False, True])
def test_build_octree_from_file(with_intensity, tmp_path):
    """generation.rs:272-287: find_bounding_box + build_octree over the PLY stream == the oracle on the same stream."""
    import point_cloud_viewer_b200 as pcv

    n = 300000
    rng = np.random.default_rng(3)
    cen = rng.random((12, 3)) * [150, 150, 15]
    P = cen[rng.integers(0, 12, n)] + rng.normal(0, 1.5, (n, 3))
    cols = {"x": P[:, 0].astype(np.float32), "y": P[:, 1].astype(np.float32), "z": P[:, 2].astype(np.float32)}
    path = str(tmp_path / "t.ply")
    write_ply(path, n, LAYOUTS[1], rng, offset=(4.1e6, 6.6e5, 4.7e6), columns={k: v.astype(np.float64) for k, v in cols.items()})
    c = pcv.Context(0, max_points_per_node=2000)
    attrs = ("color", "intensity") if with_intensity else ("color",)
    tree = c.build_octree_from_file(path, 0.001, attrs)
    x, y, z, rgb, inten = O.ply_read(path)
    bmin, bmax = O.ply_find_bounding_box(path)
    ref = O.build(x, y, z, rgb, 0.001, bmin, bmax, intensity=inten if with_intensity else None, max_points_per_node=2000)
    compare_trees(ref, tree)
    assert sum(v["num_points"] for v in tree.nodes.values()) == n
    # the directory written through the drop-in function has the reference's layout
    out = tmp_path / "oct"
    pcv.build_octree_from_file(str(out), 0.001, path, attrs, ctx=c).free()
    assert (out / "meta.pb").exists() and (out / "r.xyz").exists() and (out / "r.rgb").exists()
    assert (out / "r.intensity").exists() == with_intensity
    tree.free()
    c.close()


def test_empty_truncated_and_colourless_files(ctx, tmp_path):
    import point_cloud_viewer_b200 as pcv

    rng = np.random.default_rng(4)
    e = str(tmp_path / "e.ply")
    write_ply(e, 0, LAYOUTS[0], rng)
    pp = ctx.load_ply(e)
    assert pp.n == 0 and tuple(pp.bbox_min) == (0, 0, 0) and tuple(pp.bbox_max) == (0, 0, 0)
    t = ctx.build_octree_from_file(e, 0.001)
    assert len(t.nodes) == 0
    t.free()
    cut = str(tmp_path / "cut.ply")
    write_ply(cut, 1000, LAYOUTS[0], rng, body_cut=7)
    with pytest.raises(pcv.PcvError) as ex:
        ctx.load_ply(cut)
    assert ex.value.code == -3 and "truncated" in str(ex.value)
    nocol = str(tmp_path / "nc.ply")
    write_ply(nocol, 100, LAYOUTS[4], rng)
    with pytest.raises(pcv.PcvError) as ex:
        ctx.build_octree_from_file(nocol, 0.001)
    assert "color is mandatory" in str(ex.value)
    with pytest.raises(pcv.PcvError):
        ctx.build_octree_from_file(str(tmp_path / "l0.ply"), 0.001)  # missing file
