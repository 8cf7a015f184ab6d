"""The reference's own integration tests (src/octree/tests.rs) restated in C++ against include/pcv.hpp, run on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_octree_tests(tmp_path):
    exe = str(tmp_path / "test_octree")
    lib_dir = os.path.join(ROOT, "point_cloud_viewer_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "test_octree.cpp"), "-o", exe, "-L" + lib_dir, "-l:libpcv_b200.so",
                           "-Wl,-rpath," + lib_dir])
    d = str(tmp_path / "octree")
    os.makedirs(d)
    r = subprocess.run([exe, d], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    assert sorted(os.listdir(d)) == ["meta.pb", "r.rgb", "r.xyz", "r4.rgb", "r4.xyz"]  # r0 has no files, as in the reference
