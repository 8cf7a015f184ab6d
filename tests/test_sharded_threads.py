"""The sharding orchestration (point_cloud_viewer_b200/distributed.py: cell assignment, ShardSpec split decisions above the
shard level, collectors, top-of-tree assembly) for 2-5 ranks, k = 1..3 and assorted clouds, with the ranks as threads of one
process talking through an in-memory communicator - the same code path as the gloo test, cheap enough to sweep.  The merged
tree must equal the oracle's single build bit for bit (T11)."""
import threading

import numpy as np
import pytest

import oracle_api as O
import tb_api
import bench
import point_cloud_viewer_b200 as pcv
from point_cloud_viewer_b200 import distributed as D


class ThreadWorld:
    def __init__(self, n):
        self.n = n
        self.barrier = threading.Barrier(n)
        self.slots = [None] * n


class ThreadComm:
    """The communicator interface of distributed.TorchComm over shared memory + a barrier."""

    def __init__(self, world, rank):
        self.w, self.rank, self.world = world, rank, world.n

    def _all(self, obj):
        self.w.slots[self.rank] = obj
        self.w.barrier.wait()
        out = list(self.w.slots)
        self.w.barrier.wait()
        return out

    def all_reduce_sum_u64(self, arr):
        return np.sum([np.asarray(a, np.uint64) for a in self._all(np.asarray(arr, np.uint64))], axis=0).astype(np.uint64)

    def all_reduce_minmax(self, mn, mx):
        parts = self._all((np.asarray(mn, np.float64), np.asarray(mx, np.float64)))
        return np.min([p[0] for p in parts], axis=0), np.max([p[1] for p in parts], axis=0)

    def exchange_counts(self, send_counts):
        parts = self._all(np.asarray(send_counts, np.int64))
        return np.array([p[self.rank] for p in parts], np.int64)

    def all_to_all(self, tensor, send_counts, recv_counts, alloc=None):
        import torch

        starts = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
        parts = self._all((tensor, starts))
        return torch.cat([t[s[self.rank]:s[self.rank + 1]] for t, s in parts])

    def all_gather_objects(self, obj):
        return self._all(obj)

    def done_with(self, *tensors):
        pass

    def barrier(self):
        self.w.barrier.wait()


def _cloud(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "clusters":
        cen = rng.random((10, 3)) * [200, 200, 40]
        k = rng.integers(0, 10, n)
        P = cen[k] + rng.normal(0, 1, (n, 3)) * (rng.random((10, 1))[k] * 4 + 0.2)
        P[: n // 20] = P[0]
    elif kind == "uniform":
        P = rng.random((n, 3)) * 100
    elif kind == "corner":  # everything in one octant: most cells are empty, some ranks receive nothing
        P = rng.random((n, 3)) * [12, 12, 12]
        P[0] = [100, 100, 100]
    else:  # "plane"
        P = rng.random((n, 3)) * [100, 100, 0]
    P = P + [4.1e6, 6.6e5, 4.7e6]
    return P, rng.integers(0, 255, n * 3, dtype=np.uint8), rng.random(n).astype(np.float32)


def _run(world_size, P, rgb, inten, res, maxpts, k):
    n = len(P)
    world = ThreadWorld(world_size)
    results, errors = [None] * world_size, []
    final_bytes = [0] * world_size
    cuts = np.linspace(0, n, world_size + 1).astype(int)

    def worker(rank):
        try:
            lo, hi = cuts[rank], cuts[rank + 1]
            x, y, z = [np.ascontiguousarray(P[lo:hi, i]) for i in range(3)]
            ops = tb_api.TbOps(x, y, z, rgb[3 * lo:3 * hi].copy(), inten[lo:hi].copy(), res, P.min(0), P.max(0), maxpts)
            comm = ThreadComm(world, rank)
            tree = D.build_sharded(ops, comm, int(lo), prefix_levels=k, max_points_per_node=maxpts)
            final_bytes[rank] = bench._final_node_bytes(pcv, tree)
            results[rank] = (tree.gather_all(comm), tree.k, tree.bbox_inside)
        except BaseException as e:  # noqa: BLE001 - release the other threads
            errors.append(e)
            world.barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world_size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results[0] + (sum(final_bytes),)


CASES = [
    (2, "clusters", 30000, 300, 2, 1e-3), (3, "clusters", 30000, 300, 2, 1e-3), (4, "clusters", 40000, 200, 3, 1e-3), (5, "uniform", 20000, 150, 2, 1e-2),
    (3, "corner", 20000, 200, 2, 1e-3), (4, "plane", 20000, 100, 3, 1e-3), (2, "uniform", 5000, 100000, 2, 1e-3), (3, "clusters", 25000, 5000, 3, 1e-2),
    (2, "uniform", 3, 1, 1, 1e-3), (4, "clusters", 20000, 250, 1, 1e-9),
]


@pytest.mark.parametrize("world_size,kind,n,maxpts,k,res", CASES)
def test_sharded_build_equals_single_build(world_size, kind, n, maxpts, k, res):
    P, rgb, inten = _cloud(kind, n, world_size * 1000 + n)
    nodes, k_used, inside, final_bytes = _run(world_size, P, rgb, inten, res, maxpts, k)
    assert inside and 1 <= k_used <= k
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    ref = O.build(x, y, z, rgb.reshape(-1, 3), res, P.min(0), P.max(0), intensity=inten, max_points_per_node=maxpts)
    assert set(nodes) == set(ref.nodes), sorted(set(nodes) ^ set(ref.nodes))[:10]
    # bench.py's N > 1 roofline numerator: every final node exactly once over the ranks (sum of n (3 bpc + 3))
    assert final_bytes == sum(m["num_points"] * (3 * pcv.ENC_BYTES[m["enc"]] + 3) for m in ref.nodes.values())
    for name, m in ref.nodes.items():
        g = nodes[name]
        assert (g["num_points"], g["enc"], tuple(g["cube"])) == (m["num_points"], m["enc"], tuple(m["cube"])), name
        if m["num_points"]:
            rx, rc, ri, rs = ref.node_data(name, True)
            assert np.array_equal(rs, g["src"]), (name, "global source index order")
            assert np.array_equal(rx, g["xyz"]) and np.array_equal(rc, g["rgb"]) and np.array_equal(ri, g["intensity"]), name
