import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import point_cloud_viewer_b200 as pcv
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
kind = pcv.SYNTH_GAUSS_CLUSTERS
bmin, bmax, res = pcv.synth_bbox(kind)
ctx = pcv.Context(0, levels_per_pass=2)
x, y, z = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
rgb = torch.empty(n * 3, dtype=torch.uint8, device="cuda")
ctx.synth_points_device(kind, 1, 0, n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr())
hx, hy, hz = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(3)]
hrgb = torch.empty(n * 3, dtype=torch.uint8, pin_memory=True)
hx.copy_(x); hy.copy_(y); hz.copy_(z); hrgb.copy_(rgb)
del x, y, z, rgb
torch.cuda.empty_cache()
oxyz = torch.empty(int(n * 12 + (1 << 24)), dtype=torch.uint8, pin_memory=True)
orgb = torch.empty(n * 3, dtype=torch.uint8, pin_memory=True)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t = ctx.build_octree(hx.data_ptr(), hy.data_ptr(), hz.data_ptr(), hrgb.data_ptr(), res, bmin, bmax, n=n)
    t1 = time.perf_counter()
    t.download(xyz=oxyz.data_ptr(), rgb=orgb.data_ptr(), want_src=False)
    t2 = time.perf_counter()
    t.free()
    t3 = time.perf_counter()
    s = ctx.last_build_stats()
    print("build call %.1f ms (device build %.1f) | download %.1f ms (%.2f GB) | free %.1f ms" % ((t1 - t0) * 1e3, s["ms_total"], (t2 - t1) * 1e3, (t.xyz_bytes + 3 * n) / 1e9, (t3 - t2) * 1e3))
