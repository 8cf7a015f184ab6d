"""Scratch timing of the device-resident build (not the contract bench; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import point_cloud_viewer_b200 as pcv

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = pcv.SYNTH_GAUSS_CLUSTERS
ctx = pcv.Context(0, levels_per_pass=G)
x, y, z = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
rgb = torch.empty(n * 3, dtype=torch.uint8, device="cuda")
ctx.synth_points_device(kind, 1, 0, n, x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr())
bmin, bmax, res = pcv.synth_bbox(kind)
t0 = time.time(); mn, mx = ctx.bbox(x.data_ptr(), y.data_ptr(), z.data_ptr(), n=n, device=True); t1 = time.time()
print("bbox", mn, mx, "%.2f ms" % ((t1 - t0) * 1e3))
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
for it in range(iters):
    t0 = time.time()
    tree = ctx.build_octree(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), res, bmin, bmax, n=n, device=True)
    t1 = time.time()
    s = ctx.last_build_stats()
    print("build n=%d G=%d wall %.1f ms | total %.1f partition %.1f place %.1f ms (host plan %.1f wait %.1f) | passes %d launches %d nodes %d deepest %d | %.1f Mpts/s | algo GB/s %.0f" % (
        n, G, (t1 - t0) * 1e3, s["ms_total"], s["ms_partition"], s["ms_place"], s["ms_host_plan"], s["ms_host_wait"], s["passes"], s["kernel_launches"], s["num_nodes"], s["deepest_level"],
        n / s["ms_total"] / 1e3, s["algorithmic_bytes"] / s["ms_total"] / 1e6))
    tree.free()

ctx.set_profiling(True)
tree = ctx.build_octree(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), res, bmin, bmax, n=n, device=True)
for k, v in sorted(ctx.kernel_stats().items(), key=lambda kv: -kv[1]["ms"]):
    print("  %-22s launches %3d  %8.3f ms  %7.1f algo GB/s" % (k, v["launches"], v["ms"], v["algorithmic_bytes"] / max(v["ms"], 1e-9) / 1e6))
tree.free()
