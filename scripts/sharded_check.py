"""torchrun --nproc-per-node N scripts/sharded_check.py [points_per_rank] [max_points] [k]
T11: the N-GPU sharded build (NCCL all-to-all) equals the 1-GPU build bit for bit (node table + node contents)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import point_cloud_viewer_b200 as pcv
from point_cloud_viewer_b200 import distributed as D

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
maxpts = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
kind = pcv.SYNTH_GAUSS_CLUSTERS
bmin, bmax, res = pcv.synth_bbox(kind)
ctx = pcv.Context(local, max_points_per_node=maxpts)
x, y, z = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
rgb = torch.empty(n * 3, dtype=torch.uint8, device=dev)
ctx.synth_points_device(kind, 3, rank * n, n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr())
single = None
# fused pack+exchange over peer memory (twice: the second step reuses the IPC slab), then the staged NCCL all-to-all path
for mode in ("fused", "fused", "native", "native", "staged"):
  if mode == "staged":
      os.environ["PCV_NO_FUSED_EXCHANGE"] = "1"
  if mode == "native":  # the whole orchestration inside the C library (pcv_build_octree_sharded), Python lends the collectives
      tree = D.build_octree_sharded_native(ctx, x, y, z, rgb, None, rank * n, res, bmin, bmax, prefix_levels=k)
  else:
      tree = D.build_octree_sharded(ctx, x, y, z, rgb, None, rank * n, res, bmin, bmax, prefix_levels=k, max_points_per_node=maxpts)
  merged = tree.gather_all(D.TorchComm(dev))
  tree.free()
  if rank == 0:
      N = n * world
      X, Y, Z = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(3)]
      RGB = torch.empty(N * 3, dtype=torch.uint8, device=dev)
      ctx.synth_points_device(kind, 3, 0, N, X.data_ptr(), Y.data_ptr(), Z.data_ptr(), RGB.data_ptr())
      if single is None:
          single = ctx.build_octree(X.data_ptr(), Y.data_ptr(), Z.data_ptr(), RGB.data_ptr(), res, bmin, bmax, n=N, device=True)
      assert set(single.nodes) == set(merged), sorted(set(single.nodes) ^ set(merged))[:10]
      for name, m in single.nodes.items():
          g = merged[name]
          assert (g["num_points"], g["enc"], tuple(g["cube"])) == (m["num_points"], m["enc"], tuple(m["cube"])), name
          if m["num_points"]:
              sx, sc, si, ss = single.node_data(name)
              assert np.array_equal(ss, g["src"]), (name, "src")
              assert np.array_equal(sx, g["xyz"]) and np.array_equal(sc, g["rgb"]), name
      print("sharded (%s) == single: %d nodes, %d points, k=%d, ranks=%d OK" % (mode, len(merged), N, tree.k, world))
  dist.barrier()
dist.barrier()
dist.destroy_process_group()
