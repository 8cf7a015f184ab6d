"""point_cloud_viewer_b200/geometry.py builds the Location both the product and the oracle consume, so it needs a witness
of its own: scipy's Rotation for the quaternion algebra, numpy.linalg for the inverses, and the
textbook OpenGL frustum formula for the perspective matrix — none of which share code with geometry.py."""
import math

import numpy as np
from scipy.spatial.transform import Rotation

import point_cloud_viewer_b200.geometry as G


def test_quaternion_algebra_matches_scipy():
    rng = np.random.default_rng(11)
    for _ in range(200):
        axis, ang = rng.normal(size=3), float(rng.uniform(-math.pi, math.pi))
        q = G.quat_from_axis_angle(axis, ang)
        r = Rotation.from_rotvec(axis / np.linalg.norm(axis) * ang)
        assert np.allclose(q, r.as_quat(), atol=1e-15) or np.allclose(q, -r.as_quat(), atol=1e-15)  # scipy: (x, y, z, w) like nalgebra's (i, j, k, w)
        p = rng.normal(size=3) * 100
        assert np.allclose(G.quat_rotate(q, p), r.apply(p), rtol=0, atol=1e-12)
        q2 = G.quat_from_axis_angle(rng.normal(size=3), float(rng.uniform(-3, 3)))
        r2 = Rotation.from_quat(q2)
        assert np.allclose(G.quat_rotate(G.quat_mul(q, q2), p), (r * r2).apply(p), atol=1e-11)


def test_isometry_matrix_inverse_compose():
    rng = np.random.default_rng(12)
    for _ in range(100):
        a = G.Isometry(rng.normal(size=3) * 1e3, G.quat_from_axis_angle(rng.normal(size=3), float(rng.uniform(-3, 3))))
        b = G.Isometry(rng.normal(size=3) * 10, G.quat_from_axis_angle(rng.normal(size=3), float(rng.uniform(-3, 3))))
        A = np.eye(4)
        A[:3, :3] = Rotation.from_quat(a.q).as_matrix()
        A[:3, 3] = a.t
        assert np.allclose(a.to_homogeneous(), A, atol=1e-12)
        assert np.allclose(a.inverse().to_homogeneous(), np.linalg.inv(A), atol=1e-9)
        assert np.allclose((a * b).to_homogeneous(), A @ b.to_homogeneous(), atol=1e-9)
        p = rng.normal(size=3)
        assert np.allclose(a.transform_point(p), (A @ np.append(p, 1.0))[:3], atol=1e-9)


def test_perspective_is_the_opengl_frustum_and_its_inverse():
    rng = np.random.default_rng(13)
    for _ in range(100):
        l, b = -rng.uniform(0.1, 2.0), -rng.uniform(0.1, 2.0)
        r, t = rng.uniform(0.1, 2.0), rng.uniform(0.1, 2.0)
        n = rng.uniform(0.01, 2.0)
        f = n + rng.uniform(0.1, 500.0)
        P = G.Perspective(l, r, b, t, n, f)
        want = np.array([[2 * n / (r - l), 0, (r + l) / (r - l), 0], [0, 2 * n / (t - b), (t + b) / (t - b), 0], [0, 0, -(f + n) / (f - n), -2 * f * n / (f - n)],
                         [0, 0, -1, 0]])
        assert np.allclose(P.matrix, want, rtol=1e-15, atol=0)
        assert np.allclose(P.inverse(), np.linalg.inv(want), rtol=1e-9, atol=1e-12)
        # the 8 frustum corners map to the clip cube's corners
        for sx in (-1, 1):
            for sy in (-1, 1):
                for (zd, sz) in ((n, -1), (f, 1)):
                    x = (l if sx < 0 else r) * zd / n
                    y = (b if sy < 0 else t) * zd / n
                    h = want @ np.array([x, y, -zd, 1.0])
                    assert np.allclose(h[:3] / h[3], [sx, sy, sz], atol=1e-9)
    fov = G.Perspective.new_fov(1.3, 1.1, 0.5, 80.0).matrix
    assert math.isclose(fov[1, 1], 1.0 / math.tan(0.55), rel_tol=1e-14) and math.isclose(fov[0, 0], fov[1, 1] / 1.3, rel_tol=1e-14)


def test_frustum_location_fields():
    iso = G.Isometry((5.0, -3.0, 2.0), G.quat_from_axis_angle([0.3, -0.2, 0.9], 0.8))
    P = G.Perspective.new_fov(1.0, 1.2, 0.1, 10.0)
    loc = G.frustum(iso, P)
    cfq = np.array(loc.clip_from_query).reshape(4, 4).T  # column-major in the struct
    qfc = np.array(loc.query_from_clip).reshape(4, 4).T
    A = np.eye(4)
    A[:3, :3] = Rotation.from_quat(iso.q).as_matrix()
    A[:3, 3] = iso.t
    assert np.allclose(cfq, P.matrix @ np.linalg.inv(A), atol=1e-10)
    assert np.allclose(cfq @ qfc, np.eye(4), atol=1e-9)
    o = G.obb(iso, (1.0, 2.0, 3.0))
    inv = G.Isometry(o.obb_from_query[:3], o.obb_from_query[3:])
    assert np.allclose((iso * inv).to_homogeneous(), np.eye(4), atol=1e-12)
