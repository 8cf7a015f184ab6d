"""Scratch timing of the PLY input path: file (page cache) -> pinned ring -> H2D -> k_ply_unpack (+ fused bbox)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import point_cloud_viewer_b200 as pcv

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
path = os.path.join(d, "pcv_bench_%d.ply" % n)
rec = np.zeros(n, dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")]))
rng = np.random.default_rng(0)
for k in "xyz":
    rec[k] = rng.random(n, dtype=np.float32) * 1000
rec["red"] = np.arange(n) & 255
t0 = time.time()
with open(path, "wb") as f:
    f.write(("ply\nformat binary_little_endian 1.0\ncomment offset: 4000000 600000 4700000\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
             "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % n).encode())
    rec.tofile(f)
print("wrote %s %.2f GB in %.1f s" % (path, os.path.getsize(path) / 1e9, time.time() - t0))
ctx = pcv.Context(0)
ctx.set_profiling(True)
for it in range(3):
    t0 = time.time()
    pp = ctx.load_ply(path)
    dt = time.time() - t0
    print("load_ply: %.1f ms  %.1f Mpts/s  %.2f GB/s file bytes  bbox %s" % (dt * 1e3, n / dt / 1e6, os.path.getsize(path) / dt / 1e9, pp.bbox_min))
    pp.free()
ks = ctx.kernel_stats()["k_ply_unpack"]
print("k_ply_unpack: launches %d  %.3f ms  %.1f GB/s algorithmic" % (ks["launches"], ks["ms"], ks["algorithmic_bytes"] / ks["ms"] / 1e6))
ctx.set_profiling(False)
for it in range(2):
    t0 = time.time()
    tree = ctx.build_octree_from_file(path, 0.001)
    dt = time.time() - t0
    print("build_octree_from_file: %.1f ms  %.1f Mpts/s  nodes %d" % (dt * 1e3, n / dt / 1e6, len(tree.meta)))
    tree.free()
import oracle_api as O
m = min(n, 5_000_000)
t0 = time.time(); O.ply_read(path, 0, m); dt = time.time() - t0
print("oracle PlyIterator restatement: %.1f Mpts/s (1 thread, %d points)" % (m / dt / 1e6, m))
os.remove(path)
