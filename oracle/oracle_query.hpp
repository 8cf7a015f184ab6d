// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
//
// Culling math, SAT, node selection, filtered point iteration and X-ray accumulation, restated from
// src/math/sat.rs, src/geometry/{frustum,obb,aabb}.rs, src/octree/{mod,octree_iterator}.rs,
// src/iterator.rs and xray/src/generation.rs.  nalgebra 0.22.0 operation order (un-vendored) is
// restated from the crate's published source: matrix*vector accumulates column by column,
// `Vector / s` divides component-wise, UnitQuaternion*Vector3 = (t*w + qv x t) + p with t=(qv x p)*2.
#pragma once
#include <functional>

#include "oracle_core.hpp"

namespace orc {

inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 normalize(Vec3 v) {  // Unit::new_normalize: v / ||v||
    double n = std::sqrt(dot(v, v));
    return {v.x / n, v.y / n, v.z / n};
}

// 4x4, column-major (nalgebra storage): m[c*4+r].
struct Mat4 {
    double m[16];
    double at(int r, int c) const { return m[c * 4 + r]; }
    double& at(int r, int c) { return m[c * 4 + r]; }
};

// Matrix4::transform_point (nalgebra 0.22 geometry/transform_ops / base/cg.rs)
inline Vec3 transform_point(const Mat4& M, Vec3 p) {
    double n = 0.0;
    n = M.at(3, 0) * p.x;
    n = n + M.at(3, 1) * p.y;
    n = n + M.at(3, 2) * p.z;
    n = n + M.at(3, 3);
    double r[3];
    for (int i = 0; i < 3; ++i) {
        double a = M.at(i, 0) * p.x;
        a = M.at(i, 1) * p.y + a;
        a = M.at(i, 2) * p.z + a;
        r[i] = a + M.at(i, 3);
    }
    if (n != 0.0) return {r[0] / n, r[1] / n, r[2] / n};
    return {r[0], r[1], r[2]};
}

// m * p.to_homogeneous() ; Point3::from_homogeneous(..).unwrap()   (octree/mod.rs:103-106)
inline bool project(const Mat4& M, Vec3 p, Vec3& out) {
    double q[4];
    for (int i = 0; i < 4; ++i) {
        double a = M.at(i, 0) * p.x;
        a = M.at(i, 1) * p.y + a;
        a = M.at(i, 2) * p.z + a;
        a = M.at(i, 3) * 1.0 + a;
        q[i] = a;
    }
    if (q[3] == 0.0) return false;
    out = {q[0] / q[3], q[1] / q[3], q[2] / q[3]};
    return true;
}

// 4x4 inverse, cofactor expansion as in nalgebra's do_inverse4 (the MESA gluInvertMatrix formula).
inline bool try_inverse(const Mat4& A, Mat4& out) {
    const double* m = A.m;
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0) return false;
    double inv_det = 1.0 / det;
    for (int i = 0; i < 16; ++i) out.m[i] = inv[i] * inv_det;
    return true;
}

// Isometry3 = translation + unit quaternion (x,y,z,w)
struct Iso3 {
    Vec3 t;
    double q[4];  // i, j, k, w
};
inline Vec3 quat_rotate(const double q[4], Vec3 p) {  // UnitQuaternion * Vector3
    Vec3 qv{q[0], q[1], q[2]};
    Vec3 c = cross(qv, p);
    Vec3 t{c.x * 2.0, c.y * 2.0, c.z * 2.0};
    Vec3 cr = cross(qv, t);
    return {t.x * q[3] + cr.x + p.x, t.y * q[3] + cr.y + p.y, t.z * q[3] + cr.z + p.z};
}
inline Vec3 iso_transform_point(const Iso3& I, Vec3 p) {  // translation * (rotation * p)
    Vec3 r = quat_rotate(I.q, p);
    return {r.x + I.t.x, r.y + I.t.y, r.z + I.t.z};
}
inline Iso3 iso_inverse(const Iso3& I) {
    Iso3 r;
    r.q[0] = -I.q[0];
    r.q[1] = -I.q[1];
    r.q[2] = -I.q[2];
    r.q[3] = I.q[3];
    Vec3 nt{-I.t.x, -I.t.y, -I.t.z};
    r.t = quat_rotate(r.q, nt);
    return r;
}

enum Relation { REL_IN = 0, REL_CROSS = 1, REL_OUT = 2 };  // sat.rs:39-47

// sat.rs:174-205
inline Relation sat(const std::vector<Vec3>& axes, const Vec3* ca, int na, const Vec3* cb, int nb) {
    Relation rel = REL_IN;
    for (const Vec3& ax : axes) {
        double amin = std::numeric_limits<double>::max(), amax = std::numeric_limits<double>::lowest();
        for (int i = 0; i < na; ++i) {
            double p = dot(ca[i], ax);
            amin = std::fmin(amin, p);
            amax = std::fmax(amax, p);
        }
        double bmin = std::numeric_limits<double>::max(), bmax = std::numeric_limits<double>::lowest();
        for (int i = 0; i < nb; ++i) {
            double p = dot(cb[i], ax);
            bmin = std::fmin(bmin, p);
            bmax = std::fmax(bmax, p);
        }
        if (bmin > amax || bmax < amin) return REL_OUT;
        if (amin > bmin || bmax > amax) rel = REL_CROSS;
    }
    return rel;
}

struct Intersector {  // sat.rs:67-75
    Vec3 corners[8];
    std::vector<Vec3> edges, face_normals;
};

// sat.rs:80-103
inline std::vector<Vec3> separating_axes(const Intersector& a, const std::vector<Vec3>& other_edges,
                                         const std::vector<Vec3>& other_face_normals) {
    std::vector<Vec3> axes;
    for (auto& n : a.face_normals) axes.push_back(n);
    for (auto& n : other_face_normals) axes.push_back(n);
    for (auto& e1 : a.edges)
        for (auto& e2 : other_edges) {
            Vec3 c = normalize(cross(e1, e2));
            if (std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z)) axes.push_back(c);
        }
    return axes;
}

struct CachedAxesIntersector {  // sat.rs:154-157
    std::vector<Vec3> axes;
    Vec3 corners[8];
    Relation intersect(const Vec3* c, int n) const { return sat(axes, corners, 8, c, n); }
};

// sat.rs:111-134 (O(n^2) dedup keeping the first) and :138-143
inline CachedAxesIntersector cache_separating_axes_for_aabb(const Intersector& a) {
    std::vector<Vec3> unit{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    std::vector<Vec3> all = separating_axes(a, unit, unit);
    CachedAxesIntersector r;
    for (auto& ax1 : all) {
        bool dupe = false;
        for (auto& ax2 : r.axes) {
            Vec3 dm = ax1 - ax2, dp = ax1 + ax2;
            double d1 = dot(dm, dm), d2 = dot(dp, dp);
            if (std::fmin(d1, d2) < std::numeric_limits<double>::epsilon()) {
                dupe = true;
                break;
            }
        }
        if (!dupe) r.axes.push_back(ax1);
    }
    std::memcpy(r.corners, a.corners, sizeof(r.corners));
    return r;
}

// ----- Frustum (frustum.rs:95-166) -----
struct Frustum {
    Mat4 query_from_clip, clip_from_query;
    bool contains(Vec3 p) const {  // frustum.rs:120-125
        Vec3 c = transform_point(clip_from_query, p);
        double mn = std::fmin(std::fmin(c.x, c.y), c.z), mx = std::fmax(std::fmax(c.x, c.y), c.z);
        return mn > -1.0 && mx < 1.0;
    }
    void corners(Vec3 c[8]) const {  // frustum.rs:129-141
        int i = 0;
        for (double x : {-1.0, 1.0})
            for (double y : {-1.0, 1.0})
                for (double z : {-1.0, 1.0}) c[i++] = transform_point(query_from_clip, {x, y, z});
    }
    Intersector intersector() const {  // frustum.rs:143-166
        Intersector r;
        corners(r.corners);
        const Vec3* c = r.corners;
        r.edges = {normalize(c[4] - c[0]), normalize(c[2] - c[0]), normalize(c[1] - c[0]),
                   normalize(c[3] - c[2]), normalize(c[5] - c[4]), normalize(c[7] - c[6])};
        auto& e = r.edges;
        r.face_normals = {normalize(cross(e[0], e[1])), normalize(cross(e[0], e[2])), normalize(cross(e[0], e[3])),
                          normalize(cross(e[1], e[2])), normalize(cross(e[1], e[4]))};
        return r;
    }
};

// ----- Obb (obb.rs) -----
struct Obb {
    Iso3 query_from_obb, obb_from_query;
    Vec3 half_extent;
    bool contains(Vec3 p) const {  // obb.rs:83-90
        Vec3 q = iso_transform_point(obb_from_query, p);
        return std::fabs(q.x) <= half_extent.x && std::fabs(q.y) <= half_extent.y && std::fabs(q.z) <= half_extent.z;
    }
    Intersector intersector() const {  // obb.rs:49-78
        Intersector r;
        Vec3 h = half_extent;
        Vec3 loc[8] = {{-h.x, -h.y, -h.z}, {h.x, -h.y, -h.z}, {-h.x, h.y, -h.z}, {h.x, h.y, -h.z},
                       {-h.x, -h.y, h.z},  {h.x, -h.y, h.z},  {-h.x, h.y, h.z},  {h.x, h.y, h.z}};
        for (int i = 0; i < 8; ++i) r.corners[i] = iso_transform_point(query_from_obb, loc[i]);
        r.edges = {normalize(quat_rotate(query_from_obb.q, {1, 0, 0})), normalize(quat_rotate(query_from_obb.q, {0, 1, 0})),
                   normalize(quat_rotate(query_from_obb.q, {0, 0, 1}))};
        r.face_normals = r.edges;
        return r;
    }
};

inline Intersector aabb_intersector_generic(const Aabb& b) {  // aabb.rs:127-139
    Intersector r;
    b.corners(r.corners);
    r.edges = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    r.face_normals = r.edges;
    return r;
}

enum LocKind { LOC_ALL = 0, LOC_AABB = 1, LOC_FRUSTUM = 2, LOC_OBB = 3 };  // iterator.rs:13-20
struct Location {
    int kind = LOC_ALL;
    Aabb aabb{};
    Frustum frustum{};
    Obb obb{};
    bool contains(Vec3 p) const {
        switch (kind) {
            case LOC_AABB: return aabb.contains(p);
            case LOC_FRUSTUM: return frustum.contains(p);
            case LOC_OBB: return obb.contains(p);
            default: return true;  // math/mod.rs:157-161
        }
    }
};

// base.rs:29-46 / aabb.rs:103-111 / math/mod.rs:143-147
struct AabbIntersector {
    bool all = false;
    CachedAxesIntersector isec;
    bool intersect_aabb(const Aabb& b) const {
        if (all) return true;
        Vec3 c[8];
        b.corners(c);
        return isec.intersect(c, 8) != REL_OUT;
    }
};
inline AabbIntersector make_aabb_intersector(const Location& loc) {
    AabbIntersector r;
    switch (loc.kind) {
        case LOC_AABB:
            r.isec.axes = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            loc.aabb.corners(r.isec.corners);
            break;
        case LOC_FRUSTUM: r.isec = cache_separating_axes_for_aabb(loc.frustum.intersector()); break;
        case LOC_OBB: r.isec = cache_separating_axes_for_aabb(loc.obb.intersector()); break;
        default: r.all = true;
    }
    return r;
}

// octree/mod.rs:309-323 + octree_iterator.rs:30-43 : BFS over existing nodes.
inline std::vector<NodeId> nodes_in_location(const Octree& oct, const Location& loc) {
    AabbIntersector isec = make_aabb_intersector(loc);
    std::vector<NodeId> out;
    std::deque<NodeId> q;
    q.push_back(NodeId());
    while (!q.empty()) {
        NodeId cur = q.front();
        q.pop_front();
        auto it = oct.nodes.find(cur);
        if (it == oct.nodes.end()) {
            // `octree.nodes[&node_id]` would panic in the reference (only possible for an empty octree).
            continue;
        }
        if (isec.intersect_aabb(it->second.cube.to_aabb())) {
            for (unsigned k = 0; k < 8; ++k) {
                NodeId c = cur.child(k);
                if (oct.nodes.count(c)) q.push_back(c);
            }
            out.push_back(cur);
        }
    }
    return out;
}

// octree/mod.rs:103-139
inline double clampd(double x, double lo, double hi) { return num_clamp(x, lo, hi); }
inline double relative_size_on_screen(const Cube& cube, const Mat4& M) {
    Vec3 mn = cube.min, mx = cube.max();
    auto proj = [&](Vec3 p) {
        Vec3 q{0, 0, 0};
        project(M, p, q);  // reference unwrap()s; w==0 is a panic there
        return Vec3{clampd(q.x, -1., 1.), clampd(q.y, -1., 1.), clampd(q.z, 0., 1.)};
    };
    Aabb rv = Aabb::make(proj(mn), proj(mx));
    Vec3 ps[6] = {{mx.x, mn.y, mn.z}, {mn.x, mx.y, mn.z}, {mx.x, mx.y, mn.z},
                  {mn.x, mn.y, mx.z}, {mx.x, mn.y, mx.z}, {mn.x, mx.y, mx.z}};
    for (auto& p : ps) rv.grow(proj(p));
    Vec3 d = rv.diag();
    return d.x * d.y;
}

// Rust std BinaryHeap<OpenNode> restated (push = sift_up; pop = swap-last-to-root,
// sift_down_to_bottom choosing right when left <= right, then sift_up), octree/mod.rs:360-404.
struct OpenNode {
    NodeId id;
    Cube cube;
    Relation relation;
    double size_on_screen;
    bool empty;
};
struct RustBinaryHeap {
    std::vector<OpenNode> d;
    static bool le(const OpenNode& a, const OpenNode& b) { return a.size_on_screen <= b.size_on_screen; }
    static bool gt(const OpenNode& a, const OpenNode& b) { return a.size_on_screen > b.size_on_screen; }
    void sift_up(size_t start, size_t pos) {
        OpenNode hole = d[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (le(hole, d[parent])) break;
            d[pos] = d[parent];
            pos = parent;
        }
        d[pos] = hole;
    }
    void push(const OpenNode& n) {
        size_t old = d.size();
        d.push_back(n);
        sift_up(0, old);
    }
    bool pop(OpenNode& out) {
        if (d.empty()) return false;
        OpenNode item = d.back();
        d.pop_back();
        if (!d.empty()) {
            std::swap(item, d[0]);
            size_t end = d.size(), pos = 0;
            OpenNode hole = d[0];
            size_t child = 1;
            while (child < end) {
                size_t right = child + 1;
                if (right < end && !gt(d[child], d[right])) child = right;
                d[pos] = d[child];
                pos = child;
                child = 2 * pos + 1;
            }
            d[pos] = hole;
            sift_up(0, pos);
        }
        out = item;
        return true;
    }
};

// octree/mod.rs:228-283
inline bool get_visible_nodes(const Octree& oct, const Mat4& M, std::vector<NodeId>& visible) {
    Frustum fr;
    fr.clip_from_query = M;
    if (!try_inverse(M, fr.query_from_clip)) return false;  // "Invalid projection matrix." panic
    CachedAxesIntersector isec = cache_separating_axes_for_aabb(fr.intersector());
    RustBinaryHeap open;
    auto maybe_push = [&](Relation rel, NodeId id, const Cube& cube) {  // mod.rs:388-404
        auto it = oct.nodes.find(id);
        if (it == oct.nodes.end()) return;
        OpenNode n{id, cube, rel, relative_size_on_screen(cube, M), it->second.num_points == 0};
        open.push(n);
    };
    maybe_push(REL_CROSS, NodeId(), oct.root_cube());
    OpenNode cur;
    while (open.pop(cur)) {
        if (cur.relation == REL_CROSS) {
            for (unsigned k = 0; k < 8; ++k) {
                Cube cc = get_child_cube(cur.cube, k);
                Vec3 c[8];
                cc.to_aabb().corners(c);
                Relation r = isec.intersect(c, 8);
                if (r == REL_OUT) continue;
                maybe_push(r, cur.id.child(k), cc);
            }
        } else {
            for (unsigned k = 0; k < 8; ++k) maybe_push(REL_IN, cur.id.child(k), get_child_cube(cur.cube, k));
        }
        if (!cur.empty) visible.push_back(cur.id);
    }
    return true;
}

// iterator.rs:96-119 (FilteredIterator) for one node: decode, cull, attribute intervals, retain.
struct Interval {
    int attribute;  // 0 = intensity (the only 1-d attribute an octree stores; iterator.rs:82-91)
    double lo, hi;
};
struct QueryOut {
    std::vector<double> xyz;  // AoS like PointsBatch.position
    std::vector<uint8_t> rgb;
    std::vector<float> intensity;
    std::vector<uint64_t> src;
};
inline void query_node(const Octree& oct, NodeId id, const Location& loc, const std::vector<Interval>& filters,
                       QueryOut& out) {
    auto it = oct.files.find(id);
    if (it == oct.files.end()) return;  // num_points == 0 -> empty NodeIterator (node_iterator.rs:56-60)
    const NodeFile& f = it->second;
    size_t n = (size_t)f.num_points();
    for (size_t i = 0; i < n; ++i) {
        Point pt = node_read(f, i, oct.with_intensity);
        bool keep = loc.contains(pt.p);
        for (auto& fi : filters) {
            double v = (double)pt.intensity;
            keep = keep && (fi.lo <= v && v <= fi.hi);  // math/mod.rs:87-89
        }
        if (!keep) continue;
        out.xyz.push_back(pt.p.x);
        out.xyz.push_back(pt.p.y);
        out.xyz.push_back(pt.p.z);
        out.rgb.push_back(pt.rgb[0]);
        out.rgb.push_back(pt.rgb[1]);
        out.rgb.push_back(pt.rgb[2]);
        if (oct.with_intensity) out.intensity.push_back(pt.intensity);
        out.src.push_back(pt.src);
    }
}

// Pixels without a colour get `TRANSPARENT.to_u8()` = (255, 255, 255, 0): generation.rs:506-511, src/color.rs:154-159
// (TRANSPARENT is white with alpha 0, not zero).
inline void fill_transparent(std::vector<uint8_t>& rgba, size_t npix) {
    rgba.resize(npix * 4);
    for (size_t i = 0; i < npix; ++i) {
        rgba[i * 4 + 0] = rgba[i * 4 + 1] = rgba[i * 4 + 2] = 255;
        rgba[i * 4 + 3] = 0;
    }
}

// xray/src/generation.rs:464-513 with the XRay strategy (:159-198) and process_point_data (:108-127).
// `has_q` selects the OBB location + query_from_global transform (:471-477,493-497).
inline bool xray_tile(const Octree& oct, const Aabb& bbox, uint32_t w, uint32_t h, bool has_q, const Iso3& query_from_global,
                      std::vector<uint8_t>& rgba, std::vector<uint32_t>* zbits_out, std::vector<uint8_t>* zover_out = nullptr) {
    Location loc;
    if (has_q) {
        Iso3 global_from_query = iso_inverse(query_from_global);
        // Obb::from(bbox): Isometry(center, identity), half = diag*0.5  (obb.rs:19-26; aabb.rs:50-52 center = (min+max)*0.5 via nalgebra::center)
        Vec3 c{(bbox.mins.x + bbox.maxs.x) * 0.5, (bbox.mins.y + bbox.maxs.y) * 0.5, (bbox.mins.z + bbox.maxs.z) * 0.5};
        Vec3 d = bbox.diag();
        Obb o;
        o.half_extent = {d.x * 0.5, d.y * 0.5, d.z * 0.5};
        // transformed(): Obb::new(global_from_query * query_from_obb, half)  (obb.rs:43-45)
        Iso3 qfo;  // product with identity rotation: translation = g.t + g.rot*c ; rotation = g.q * identity
        Vec3 sh = quat_rotate(global_from_query.q, c);
        qfo.t = {global_from_query.t.x + sh.x, global_from_query.t.y + sh.y, global_from_query.t.z + sh.z};
        // quaternion product g.q * (0,0,0,1): w = gw*1 - gi*0 - gj*0 - gk*0 ; i = gw*0 + gi*1 + gj*0 - gk*0 ...
        const double* g = global_from_query.q;
        qfo.q[3] = g[3] * 1.0 - g[0] * 0.0 - g[1] * 0.0 - g[2] * 0.0;
        qfo.q[0] = g[3] * 0.0 + g[0] * 1.0 + g[1] * 0.0 - g[2] * 0.0;
        qfo.q[1] = g[3] * 0.0 - g[0] * 0.0 + g[1] * 1.0 + g[2] * 0.0;
        qfo.q[2] = g[3] * 0.0 + g[0] * 0.0 - g[1] * 0.0 + g[2] * 1.0;
        o.query_from_obb = qfo;
        o.obb_from_query = iso_inverse(qfo);
        loc.kind = LOC_OBB;
        loc.obb = o;
    } else {
        loc.kind = LOC_AABB;
        loc.aabb = bbox;
    }
    std::vector<uint32_t> zbits((size_t)w * h * 32, 0u);  // 1024 z buckets per pixel (generation.rs:31)
    std::vector<uint8_t> zover((size_t)w * h, 0);  // bucket index >= 1024 (only reachable through fp noise / closed OBB faces)
    bool seen_any = false;
    std::vector<Interval> nofilter;
    Vec3 mn = bbox.mins, dg = bbox.diag();
    for (NodeId id : nodes_in_location(oct, loc)) {
        QueryOut q;
        query_node(oct, id, loc, nofilter, q);
        size_t n = q.src.size();
        for (size_t i = 0; i < n; ++i) {
            seen_any = true;
            Vec3 p{q.xyz[3 * i], q.xyz[3 * i + 1], q.xyz[3 * i + 2]};
            if (has_q) p = iso_transform_point(query_from_global, p);
            uint32_t x = rust_f64_as_u32(((p.x - mn.x) / dg.x) * (double)w);
            uint32_t y = rust_f64_as_u32((1. - ((p.y - mn.y) / dg.y)) * (double)h);
            uint32_t z = rust_f64_as_u32(((p.z - mn.z) / dg.z) * 1024.);
            // The reference keys a hash map by (x,y) and a hash set by z; get_pixel_color is only asked for
            // x<w, y<h, so out-of-image keys are never observed; z >= 1024 stays a distinct set member.
            if (x < w && y < h) {
                if (z < 1024)
                    zbits[((size_t)y * w + x) * 32 + (z >> 5)] |= 1u << (z & 31);
                else
                    zover[(size_t)y * w + x] = 1;  // z == 1024: p.z on the closed OBB face; one extra set member
            }
        }
    }
    fill_transparent(rgba, (size_t)w * h);
    if (!seen_any) return false;
    double max_sat = std::log(1024.);
    for (size_t px = 0; px < (size_t)w * h; ++px) {
        uint32_t cnt = 0;
        for (int k = 0; k < 32; ++k) cnt += (uint32_t)__builtin_popcount(zbits[px * 32 + k]);
        cnt += zover[px];
        if (cnt == 0) continue;  // transparent
        double saturation = std::log((double)cnt) / max_sat;
        uint32_t v = rust_f64_as_u32((1. - saturation) * 255.);
        uint8_t g = (uint8_t)(v > 255 ? 255 : v);
        rgba[px * 4 + 0] = g;
        rgba[px * 4 + 1] = g;
        rgba[px * 4 + 2] = g;
        rgba[px * 4 + 3] = 255;
    }
    if (zbits_out) *zbits_out = std::move(zbits);
    if (zover_out) *zover_out = std::move(zover);
    return true;
}

// ---- the other colouring strategies of the X-ray tiles (xray/src/generation.rs:200-405), binning = None ------------------
// Points arrive in the canonical order (nodes in BFS order, points in node order).  The reference receives its batches
// from several worker threads in unspecified order and accumulates in f32 / f64 without compensation, so its own output
// is reproducible only up to rounding; tests compare with a tolerance of one grey level.  With `Binning = None` every
// column has the single bin 0, so the outer mean over bins is `(0 + mean) / 1`.
//   mode 1  PointColorColoringStrategy  (:294-363): per column sum of Color<f32> (u8 / 255), mean, to_u8 (`as u8`)
//   mode 2  IntensityColoringStrategy   (:210-290): per column mean intensity, clamped to [min, max],
//                                                   brighten = ln(mean - min) / ln(max - min)  (f32), grey = to_u8.
//           The reference `return`s from the whole batch at the first negative intensity (:245-247), which makes the
//           result depend on batch boundaries; restated as "points with negative intensity are skipped".
//   mode 3  HeightStddevColoringStrategy (:365-405): stats::OnlineStats per column over the (transformed) z, colour =
//           colormap(clamp(stddev as f32, 0, max_stddev) / max_stddev); colormap 0 = Jet, 1 = Monochrome(PURPLISH)
//           (xray/src/colormap.rs).  OnlineStats (crate `stats`, un-vendored) restated from memory as the Welford update
//           with population variance: parity unpinned.
inline uint8_t f32_to_u8(float v) {  // Color<f32>::to_u8: (v * 255.) as u8
    const float s = v * 255.f;
    if (!(s == s) || s <= 0.f) return 0;
    if (s >= 255.f) return 255;
    return (uint8_t)s;
}
inline float jet_base(float val) {
    auto interp = [](float v, float y0, float x0, float y1, float x1) { return (v - x0) * (y1 - y0) / (x1 - x0) + y0; };
    if (val <= -0.75f) return 0.f;
    if (val <= -0.25f) return interp(val, 0.0f, -0.75f, 1.0f, -0.25f);
    if (val <= 0.25f) return 1.0f;
    if (val <= 0.75f) return interp(val, 1.0f, 0.25f, 0.0f, 0.75f);
    return 0.0f;
}
inline void colormap_u8(int colormap, float val, uint8_t out[4]) {
    if (colormap == 0) {  // Jet
        out[0] = f32_to_u8(jet_base(val - 0.5f));
        out[1] = f32_to_u8(jet_base(val));
        out[2] = f32_to_u8(jet_base(val + 0.5f));
    } else {  // Monochrome(PURPLISH = 0.8, 0.8, 1.0)
        out[0] = f32_to_u8((1.0f - val) * 0.8f);
        out[1] = f32_to_u8((1.0f - val) * 0.8f);
        out[2] = f32_to_u8((1.0f - val) * 1.0f);
    }
    out[3] = f32_to_u8(1.0f);
}

inline Location xray_location(const Aabb& bbox, bool has_q, const Iso3& query_from_global) {
    Location loc;
    if (has_q) {  // as in xray_tile above
        Iso3 global_from_query = iso_inverse(query_from_global);
        Vec3 c{(bbox.mins.x + bbox.maxs.x) * 0.5, (bbox.mins.y + bbox.maxs.y) * 0.5, (bbox.mins.z + bbox.maxs.z) * 0.5};
        Vec3 d = bbox.diag();
        Obb o;
        o.half_extent = {d.x * 0.5, d.y * 0.5, d.z * 0.5};
        Iso3 qfo;
        Vec3 sh = quat_rotate(global_from_query.q, c);
        qfo.t = {global_from_query.t.x + sh.x, global_from_query.t.y + sh.y, global_from_query.t.z + sh.z};
        const double* g = global_from_query.q;
        qfo.q[3] = g[3] * 1.0 - g[0] * 0.0 - g[1] * 0.0 - g[2] * 0.0;
        qfo.q[0] = g[3] * 0.0 + g[0] * 1.0 + g[1] * 0.0 - g[2] * 0.0;
        qfo.q[1] = g[3] * 0.0 - g[0] * 0.0 + g[1] * 1.0 + g[2] * 0.0;
        qfo.q[2] = g[3] * 0.0 + g[0] * 0.0 - g[1] * 0.0 + g[2] * 1.0;
        o.query_from_obb = qfo;
        o.obb_from_query = iso_inverse(qfo);
        loc.kind = LOC_OBB;
        loc.obb = o;
    } else {
        loc.kind = LOC_AABB;
        loc.aabb = bbox;
    }
    return loc;
}

inline bool xray_tile_attr(const Octree& oct, const Aabb& bbox, uint32_t w, uint32_t h, bool has_q, const Iso3& query_from_global, int mode,
                           float p0, float p1, int colormap, std::vector<uint8_t>& rgba) {
    const Location loc = xray_location(bbox, has_q, query_from_global);
    const size_t npix = (size_t)w * h;
    std::vector<float> sum(npix * 4, 0.f);
    std::vector<uint64_t> count(npix, 0);
    std::vector<double> mean(npix, 0.0), variance(npix, 0.0);  // OnlineStats
    bool seen_any = false;
    std::vector<Interval> nofilter;
    Vec3 mn = bbox.mins, dg = bbox.diag();
    for (NodeId id : nodes_in_location(oct, loc)) {
        QueryOut q;
        query_node(oct, id, loc, nofilter, q);
        const size_t n = q.src.size();
        for (size_t i = 0; i < n; ++i) {
            seen_any = true;
            Vec3 p{q.xyz[3 * i], q.xyz[3 * i + 1], q.xyz[3 * i + 2]};
            if (has_q) p = iso_transform_point(query_from_global, p);
            const uint32_t x = rust_f64_as_u32(((p.x - mn.x) / dg.x) * (double)w);
            const uint32_t y = rust_f64_as_u32((1. - ((p.y - mn.y) / dg.y)) * (double)h);
            if (!(x < w && y < h)) continue;
            const size_t px = (size_t)y * w + x;
            if (mode == 1) {
                sum[px * 4 + 0] += (float)q.rgb[3 * i] / 255.f;
                sum[px * 4 + 1] += (float)q.rgb[3 * i + 1] / 255.f;
                sum[px * 4 + 2] += (float)q.rgb[3 * i + 2] / 255.f;
                sum[px * 4 + 3] += 255.f / 255.f;
                count[px]++;
            } else if (mode == 2) {
                const float v = q.intensity.empty() ? 0.f : q.intensity[i];
                if (v < 0.f) continue;
                sum[px * 4] += v;
                count[px]++;
            } else {
                const double sample = p.z, oldmean = mean[px], prevq = variance[px] * (double)count[px];
                count[px]++;
                mean[px] += (sample - oldmean) / (double)count[px];
                variance[px] = (prevq + (sample - oldmean) * (sample - mean[px])) / (double)count[px];
            }
        }
    }
    fill_transparent(rgba, npix);
    if (!seen_any) return false;
    for (size_t px = 0; px < npix; ++px) {
        if (count[px] == 0) continue;
        uint8_t* o = &rgba[px * 4];
        if (mode == 1) {
            for (int k = 0; k < 4; ++k) o[k] = f32_to_u8((0.f + sum[px * 4 + k] / (float)count[px]) / 1.f);
        } else if (mode == 2) {
            float m = (0.f + sum[px * 4] / (float)count[px]) / 1.f;
            m = std::fmin(std::fmax(m, p0), p1);  // f32::max / f32::min
            const float brighten = std::log(m - p0) / std::log(p1 - p0);
            o[0] = o[1] = o[2] = f32_to_u8(brighten);
            o[3] = f32_to_u8(1.f);
        } else {
            float sd = (float)std::sqrt(variance[px]);
            sd = sd < 0.f ? 0.f : (sd > p0 ? p0 : sd);  // num::clamp
            colormap_u8(colormap, sd / p0, o);
        }
    }
    return true;
}

}  // namespace orc
