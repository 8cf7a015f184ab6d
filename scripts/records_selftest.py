"""Single-GPU self-test of the record-exchange path (nranks = 1): ingest -> exchange into a local slab -> build from records;
the nodes of levels >= k must equal the plain single-GPU build.  Run under compute-sanitizer for memory checks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import point_cloud_viewer_b200 as pcv
from point_cloud_viewer_b200 import distributed as D

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000
maxpts = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
k = 2
dev = torch.device("cuda", 0)
kind = pcv.SYNTH_GAUSS_CLUSTERS
bmin, bmax, res = pcv.synth_bbox(kind)
ctx = pcv.Context(0, max_points_per_node=maxpts)
x, y, z = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
rgb = torch.empty(n * 3, dtype=torch.uint8, device=dev)
ctx.synth_points_device(kind, 3, 0, n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr())
counts, send = ctx.shard_ingest(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), None, n, res, bmin, bmax, k)
assert int(counts.sum()) == n, (counts.sum(), n)
wide, gl = ctx.shard_send_info(send)
cap = n + 4096
rec = torch.zeros(cap * (32 if wide else 16) + 256, dtype=torch.uint8, device=dev)
col = torch.zeros(cap + 64, dtype=torch.int32, device=dev)
dig = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
c2r = np.zeros(8 ** k, np.int32)
got = ctx.shard_exchange(send, k, c2r, 1, [0], [rec.data_ptr()], [col.data_ptr()] if wide else None, [dig.data_ptr()])
assert int(got[0]) == n
torch.cuda.synchronize()
pc = D.concat_counts(D.level_counts(counts, k))
local = ctx.build_octree_from_records(rec.data_ptr(), col.data_ptr() if wide else None, dig.data_ptr(), None, n, res, bmin, bmax, k, pc)
single = ctx.build_octree(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), res, bmin, bmax, n=n, device=True)
bad = 0
for name, m in single.nodes.items():
    if m["level"] < k:
        continue
    g = local.nodes.get(name)
    assert g is not None and (g["num_points"], g["enc"], tuple(g["cube"])) == (m["num_points"], m["enc"], tuple(m["cube"])), name
    if m["num_points"]:
        a, b = single.node_data(name), local.node_data(name)
        assert np.array_equal(a[3], b[3]), (name, "src")
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), name
print("records self-test OK: %d nodes at level >= %d, %d points" % (sum(1 for m in single.nodes.values() if m["level"] >= k), k, n))
ctx.shard_send_free(send)
