import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import point_cloud_viewer_b200 as pcv
from point_cloud_viewer_b200 import distributed as D
n, maxpts, k = 100000, 5000, 2
dev = torch.device("cuda", 0)
kind = pcv.SYNTH_GAUSS_CLUSTERS
bmin, bmax, res = pcv.synth_bbox(kind)
ctx = pcv.Context(0, max_points_per_node=maxpts)
x, y, z = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
rgb = torch.empty(n * 3, dtype=torch.uint8, device=dev)
ctx.synth_points_device(kind, 3, 0, n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr())
t = ctx.build_octree(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), res, bmin, bmax, n=n, device=True)
print("plain ok", t.num_nodes, flush=True)
counts = ctx.prefix_histogram_device(x.data_ptr(), y.data_ptr(), z.data_ptr(), n, res, bmin, bmax, k)
pc = D.concat_counts(D.level_counts(counts, k))
print("counts", int(counts.sum()), flush=True)
try:
    t2 = ctx.build_octree_sharded_device_soa(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), None, n, res, bmin, bmax, k, pc)
    print("raw sharded ok", t2.num_nodes, flush=True)
except Exception as e:
    print("raw sharded FAILED", e, flush=True)
