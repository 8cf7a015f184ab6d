"""The oracle's two build variants and its copy of the benchmark's input generators.

  * faithful: node contents round-trip through files between the steps, as the reference does through its output directory
    (generation.rs:39-126,195-253) - must leave exactly the directory the in-memory variant writes;
  * the generators of include/pcv_synth.h, compiled into the oracle for the CPU reference arm of bench.py, produce the
    same bits as the product library's host entry point (and, -m gpu, as the device kernel: tests/test_config2_parity_gpu.py).
"""
import filecmp
import os

import numpy as np

import oracle_api as O


def test_faithful_variant_leaves_the_same_directory(tmp_path):
    x, y, z, rgb = O.synth_points(O.SYNTH_GAUSS_CLUSTERS, 1, 0, 400_000, num_threads=3)
    bmin, bmax, res = O.synth_bbox(O.SYNTH_GAUSS_CLUSTERS)
    inten = (np.arange(400_000) % 1000).astype(np.float32)
    for with_i in (False, True):
        a, b = str(tmp_path / ("a%d" % with_i)), str(tmp_path / ("b%d" % with_i))
        os.makedirs(a)
        os.makedirs(b)
        ref = O.build(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, intensity=inten if with_i else None, max_points_per_node=3000, num_threads=4)
        ref.write_dir(a)
        secs, nn = O.build_faithful(x, y, z, rgb.reshape(-1, 3), res, bmin, bmax, b, intensity=inten if with_i else None, max_points_per_node=3000, num_threads=4)
        assert secs > 0 and nn == len(ref.nodes)
        fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
        assert fa == fb and any(f.endswith(".intensity") for f in fa) == with_i
        for f in fa:
            if f != "meta.pb":  # node order inside meta.pb follows a hash map in the reference; compared as a set below
                assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f
        ma, mb = O.load_dir(a), O.load_dir(b)
        assert {k: (v["num_points"], v["enc"]) for k, v in ma.nodes.items()} == {k: (v["num_points"], v["enc"]) for k, v in mb.nodes.items()}


def test_oracle_generators_equal_the_library_host_generators():
    import point_cloud_viewer_b200 as pcv  # host entry point only: no GPU needed

    for kind in (O.SYNTH_SLAB_ECEF, O.SYNTH_GAUSS_CLUSTERS):
        for first in (0, (1 << 20) - 1000):  # the second range crosses into the identical-point blocks
            a = O.synth_points(kind, 7, first, 50_000, num_threads=5)
            b = pcv.synth_points_host(kind, 7, first, 50_000)
            for u, v in zip(a, b):
                assert np.array_equal(u, v)
        mn, mx, res = O.synth_bbox(kind)
        mn2, mx2, res2 = pcv.synth_bbox(kind)
        assert np.array_equal(mn, mn2) and np.array_equal(mx, mx2) and res == res2
