import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_api as O
import point_cloud_viewer_b200 as pcv
from parity import compare_trees
rng = np.random.default_rng(23)
n = 80000
P0 = rng.random((n, 3)) * 100.0 + [4.1e6, 6.6e5, 4.7e6]
bmin, bmax = P0.min(0).copy(), P0.max(0).copy()
rgb = rng.integers(0, 255, n * 3, dtype=np.uint8)
WILD = {"huge": (100, [1e200, 6.6e5, 4.7e6]), "-1e300": (5000, [4.1e6 + 5, -1e300, 4.7e6 + 5]), "inf": (9000, [4.1e6 + 5, 6.6e5 + 5, np.inf]),
        "-inf": (12000, [-np.inf, 6.6e5 + 5, 4.7e6 + 5]), "nan": (20000, [np.nan, 6.6e5 + 1, 4.7e6 + 1]), "2^399": (20001, [4.1e6 + 5, 6.6e5 + 5, 2.0 ** 399]),
        "zero": (30000, [0.0, 0.0, 0.0])}
def run(P, label):
    x, y, z = [np.ascontiguousarray(P[:, i]) for i in range(3)]
    for maxpts, res, G in ((400, 1e-4, 2), (3000, 1e-9, 3), (400, 1e-3, 1)):
        ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, max_points_per_node=maxpts)
        c = pcv.Context(0, max_points_per_node=maxpts, levels_per_pass=G)
        t = c.build_octree(x, y, z, rgb, res, bmin, bmax)
        try:
            compare_trees(ref, t); print(label, (maxpts, res, G), "OK")
        except AssertionError as e:
            print(label, (maxpts, res, G), "FAIL", str(e)[:120])
            for name in ref.nodes:
                if ref.nodes[name]["num_points"] == 0: continue
                rd = ref.node_data(name); td = t.node_data(name)
                rs = rd[3]; ts = np.asarray(td[3], np.uint64)
                w = [int(np.nonzero(a_ == 20000)[0][0]) if (a_ == 20000).any() else -1 for a_ in (rs, ts)]
                if not np.array_equal(rs, ts) or w[0] >= 0 or w[1] >= 0:
                    d = np.nonzero(rs != ts)[0]
                    print("   node", name, "n", len(rs), "ndiff", len(d), "first diffs at", d[:4], "ref", rs[d[:4]], "got", ts[d[:4]], "enc", ref.nodes[name]["enc"], "nan point at ref/got", w)
                    if w[0] >= 0 and w[1] >= 0:
                        bpc = {1: 1, 2: 2, 3: 4, 4: 8}[ref.nodes[name]["enc"]]
                        print("      nan point codes ref", rd[0][3 * bpc * w[0]: 3 * bpc * (w[0] + 1)], "got", td[0][3 * bpc * w[1]: 3 * bpc * (w[1] + 1)])
        t.free(); c.close()
P = P0.copy(); P[20000] = WILD["nan"][1]; run(P, "nan")
