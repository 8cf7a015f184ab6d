"""scripts/xray_pyramid_bench.py (bench.py's last leg, run in a child process on the GPU box) with the GPU objects replaced by
oracle-backed stand-ins: checks the script's own logic - key names, byte accounting, JSON serialisability - on a machine
without a GPU.  (The stand-ins return the oracle's results, so the parity verdicts are trivially true here.)"""
import importlib.util
import json
import os
import types

import numpy as np

import oracle_api as O
import s2_api as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeTree:
    def __init__(self, ref):
        self.ref = ref

    def xray_quadtree(self, T, px, on_tile=None, keep_tiles=True, background=(255, 255, 255, 255), **kw):
        info, tiles = self.ref.xray_quadtree(T, px, background=background)
        if on_tile is not None:
            for (l, i), img in sorted(tiles.items(), key=lambda kv: (-kv[0][0], kv[0][1])):
                on_tile(l, i, img)
        info = dict(info, num_leaves=sum(1 for k in tiles if k[0] == info["deepest_level"]), ms_leaves=1.0, ms_parents=0.5, kernel_launches=7, leaf_points=123)
        return info, (tiles if keep_tiles else {})

    def free(self):
        pass


class FakeCloud:
    def __init__(self, P, level):
        r = S.split(P, level)
        self.cell_ids, self.cell_counts, self.order = r["ids"], r["counts"], r["order"]
        self.bbox_min, self.bbox_max = r["bmin"], r["bmax"]
        self.num_cells, self.num_points = len(r["ids"]), len(P)

    def build_stats(self):
        return dict(ms_device=2.0, kernel_launches=12, algorithmic_bytes=2 * self.num_points * 27)

    def query_union(self, u):
        return dict(src=self.order)

    def free(self):
        pass


class FakeBuf:
    def __init__(self, n):
        self.ptr = n

    def free(self):
        pass


def test_child_script_logic(monkeypatch, capsys):
    import point_cloud_viewer_b200 as pcv

    spec = importlib.util.spec_from_file_location("xray_pyramid_bench", os.path.join(ROOT, "scripts", "xray_pyramid_bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    host = {}

    class FakeCtx:
        def __init__(self, device=0):
            pass

        def build_octree(self, x, y, z, rgb, res, bmin, bmax):
            return FakeTree(O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax))

        def device_buffer(self, shape, typestr):
            return FakeBuf(int(np.prod(shape)))

        def synth_points_device(self, kind, seed, first, n, *ptrs):
            host["pts"] = pcv.synth_points_host(kind, seed, first, n)

        def build_s2_cloud(self, x, y, z, rgb=None, intensity=None, split_level=20, n=None, device=False):
            if device:
                x, y, z, _ = host["pts"]
            return FakeCloud(np.stack([x, y, z], 1), split_level)

        def close(self):
            pass

    monkeypatch.setattr(pcv, "Context", FakeCtx)
    monkeypatch.setattr("sys.argv", ["xray_pyramid_bench.py", "--points", "200000", "--tile-px", "64", "--parity-px", "32", "--peak", "6570.3", "--s2-points", "50000"])
    mod.main()
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    q, s2 = out["xray_quadtree"], out["s2_cloud"]
    assert "error" not in q and "error" not in s2, (q, s2)
    assert q["parity_check"]["equal"] is True and q["parity_check"]["tile_px"] == 32 and q["cpu_baseline"]["kind"] == "port"
    qt = q["quadtree"]
    assert qt["tile_px"] == 64 and qt["nodes"] == qt["leaves"] + qt["parents"] and qt["parents_roofline"]["algorithmic_bytes"] > 0
    # every parent: its existing children + itself, 64 x 64 x 4 bytes each
    assert qt["parents_roofline"]["algorithmic_bytes"] % (64 * 64 * 4) == 0 and 0 < qt["parents_roofline"]["frac"] < 1
    assert s2["points"] == 50000 and s2["parity_check"]["equal"] is True and s2["roofline"]["algorithmic_bytes"] == 2 * 50000 * 27
    assert s2["cells"] > 0 and s2["cpu_baseline"]["value"] > 0 and s2["Mpoints_per_s"] == 50000 / (2.0 * 1e3)
