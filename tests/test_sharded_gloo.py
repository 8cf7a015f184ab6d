"""N > 1 path on CPU: world_size-2 gloo run of the sharding orchestration (point_cloud_viewer_b200/distributed.py) with the
test-only sequential stand-ins for the kernels; the merged tree must equal the oracle's single build bit for bit (T11)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data(n, seed):
    rng = np.random.default_rng(seed)
    cen = rng.random((12, 3)) * [200, 200, 40]
    k = rng.integers(0, 12, n)
    P = cen[k] + rng.normal(0, 1, (n, 3)) * (rng.random((12, 1))[k] * 4 + 0.2)
    P[: n // 20] = P[0]
    P += [4.1e6, 6.6e5, 4.7e6]
    rgb = rng.integers(0, 255, n * 3, dtype=np.uint8)
    inten = rng.random(n).astype(np.float32)
    return P, rgb, inten


def _worker(rank, world, port, n, maxpts, k, res, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tb_api
    from point_cloud_viewer_b200 import distributed as D

    P, rgb, inten = _data(n, 5)
    per = n // world
    lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
    x, y, z = [np.ascontiguousarray(P[lo:hi, i]) for i in range(3)]
    ops = tb_api.TbOps(x, y, z, rgb[3 * lo:3 * hi].copy(), inten[lo:hi].copy(), res, P.min(0), P.max(0), maxpts)
    tree = D.build_sharded(ops, D.TorchComm(torch.device("cpu")), lo, prefix_levels=k, max_points_per_node=maxpts)
    assert tree.bbox_inside
    merged = tree.gather_all(D.TorchComm(torch.device("cpu")))
    if rank == 0:
        import pickle

        with open(out_path, "wb") as f:
            pickle.dump(dict(nodes=merged, k=tree.k, recv=tree.recv_points), f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,maxpts,k,res", [(60000, 400, 2, 0.001), (40000, 300, 1, 0.001), (50000, 30000, 2, 0.01)])
def test_two_rank_sharded_build_equals_single_build(tmp_path, n, maxpts, k, res):
    import pickle

    import oracle_api as O

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "merged.pkl")
    mp.spawn(_worker, args=(2, port, n, maxpts, k, res, out), nprocs=2, join=True)
    got = pickle.load(open(out, "rb"))
    P, rgb, inten = _data(n, 5)
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, P.min(0), P.max(0), intensity=inten, max_points_per_node=maxpts)
    nodes = got["nodes"]
    assert set(nodes) == set(ref.nodes), sorted(set(nodes) ^ set(ref.nodes))[:10]
    for name, m in ref.nodes.items():
        g = nodes[name]
        assert (g["num_points"], g["enc"], tuple(g["cube"])) == (m["num_points"], m["enc"], tuple(m["cube"])), name
        if m["num_points"]:
            rx, rc, ri, rs = ref.node_data(name, True)
            assert np.array_equal(rs, g["src"]), (name, "global source index order")
            assert np.array_equal(rx, g["xyz"]) and np.array_equal(rc, g["rgb"]) and np.array_equal(ri, g["intensity"]), name
    assert got["k"] <= k
