"""pcv_build_octree_sharded (the multi-GPU build as one C call taking the collectives as callbacks) with a one-rank
communicator: the result - every node of the local tree from level k down plus the assembled top - must equal the plain
single-GPU build bit for bit, provenance included.  The 2-rank run of the same call is scripts/sharded_check.py ("native")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class SoloComm:
    world, rank = 1, 0

    def __init__(self, device):
        self.device = device

    def all_reduce_sum_u64(self, a):
        return np.asarray(a, np.uint64)

    def all_gather_bytes(self, buf, nbytes):
        out = np.zeros((1, nbytes), np.uint8)
        out[0, : len(buf)] = buf
        return out

    def all_gather_counts(self, c):
        return np.asarray(c, np.int64)[None, :]

    def all_gather_objects(self, o):
        return [o]

    def all_to_all(self, t, sc, rc, alloc=None):
        return t

    def barrier(self):
        import torch

        torch.cuda.synchronize()


@pytest.mark.parametrize("kind_name,k,with_int,fused,res_override", [
    ("SYNTH_GAUSS_CLUSTERS", 2, False, True, None), ("SYNTH_SLAB_ECEF", 2, True, True, None), ("SYNTH_GAUSS_CLUSTERS", 1, False, True, None),
    ("SYNTH_GAUSS_CLUSTERS", 2, True, False, None),
    # resolution 1e-6 over a 1024 m cube: the upper levels are Float64 encoded, i.e. wide (32-byte) records and separate colours
    ("SYNTH_GAUSS_CLUSTERS", 2, True, True, 1e-6), ("SYNTH_GAUSS_CLUSTERS", 2, False, False, 1e-6)])
def test_one_rank_native_sharded_build_equals_plain_build(kind_name, k, with_int, fused, res_override, monkeypatch):
    import torch

    import point_cloud_viewer_b200 as pcv
    from point_cloud_viewer_b200 import distributed as D

    kind = getattr(pcv, kind_name)
    n, maxpts = 600_000, 5000
    dev = torch.device("cuda", 0)
    bmin, bmax, res = pcv.synth_bbox(kind)
    if res_override:
        res = res_override
    ctx = pcv.Context(0, max_points_per_node=maxpts)
    x, y, z = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
    rgb = torch.empty(n * 3, dtype=torch.uint8, device=dev)
    ctx.synth_points_device(kind, 11, 0, n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr())
    inten = torch.rand(n, dtype=torch.float32, device=dev) if with_int else None
    if not fused:  # the exchange of ingested records + the owner's full build (the fall-back of the fused exchange pass)
        monkeypatch.setenv("PCV_NO_FUSED_PASS", "1")
    single = ctx.build_octree(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), res, bmin, bmax, intensity=inten.data_ptr() if with_int else None, n=n, device=True)
    comm = SoloComm(dev)
    for _ in range(2):  # the second call reuses the cached slab
        tree = D.build_octree_sharded_native(ctx, x, y, z, rgb, inten, 0, res, bmin, bmax, prefix_levels=k, comm=comm)
        assert (ctx.shard_send_cells(tree.send_handle[1]) is not None) == (fused and tree.k == 2)
        merged = tree.gather_all(comm)
        assert set(single.nodes) == set(merged)
        for name, m in single.nodes.items():
            g = merged[name]
            assert (g["num_points"], g["enc"], tuple(g["cube"])) == (m["num_points"], m["enc"], tuple(m["cube"])), name
            if m["num_points"]:
                sx, sc, si, ss = single.node_data(name)
                assert np.array_equal(sx, g["xyz"]) and np.array_equal(sc, g["rgb"]) and np.array_equal(ss, g["src"]), name
                if with_int:
                    assert np.array_equal(si, g["intensity"]), name
        cs = tree.c_comm
        tree.free()
    ctx.sharded_release(cs)
    single.free()
    ctx.close()
