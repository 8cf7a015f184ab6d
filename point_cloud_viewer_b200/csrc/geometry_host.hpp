// geometry_host.hpp — per-query setup of the culling geometry (host, runs once per location):
// corners, edges, face normals and the cached, de-duplicated separating axes that the CUDA SAT
// kernel then applies to every octree node.
//   Frustum corners/edges/normals   src/geometry/frustum.rs:129-166
//   Obb corners/edges               src/geometry/obb.rs:49-78
//   Aabb fast path (3 unit axes)    src/geometry/aabb.rs:103-111
//   axis generation + O(n^2) dedup  src/math/sat.rs:80-143
//   Matrix4::transform_point        nalgebra 0.22 (column-by-column accumulation, divide by w if != 0)
#pragma once
#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/pcv.h"

namespace pcv {

struct V3 {
    double x, y, z;
};
#if defined(__CUDACC__)
#define PCV_GHD __host__ __device__ inline
#else
#define PCV_GHD inline
#endif

PCV_GHD V3 v3sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PCV_GHD V3 v3add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PCV_GHD double v3dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PCV_GHD V3 v3cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// m: column-major 4x4.  (M3x3 * p + t) / (r3 . p + m33) unless the normaliser is exactly zero.
PCV_GHD V3 mat4_transform_point(const double* m, V3 p) {
    double n = m[3] * p.x;
    n = n + m[7] * p.y;
    n = n + m[11] * p.z;
    n = n + m[15];
    double r[3];
    for (int i = 0; i < 3; ++i) {
        double a = m[i] * p.x;
        a = m[4 + i] * p.y + a;
        a = m[8 + i] * p.z + a;
        r[i] = a + m[12 + i];
    }
    if (n != 0.0) return V3{r[0] / n, r[1] / n, r[2] / n};
    return V3{r[0], r[1], r[2]};
}

// UnitQuaternion * Vector3 and Isometry3 * Point3 (iso = tx,ty,tz,qi,qj,qk,qw)
PCV_GHD V3 quat_rot(const double* iso7, V3 p) {
    const V3 qv{iso7[3], iso7[4], iso7[5]};
    V3 t = v3cross(qv, p);
    t = V3{t.x * 2.0, t.y * 2.0, t.z * 2.0};
    const V3 c = v3cross(qv, t);
    const double w = iso7[6];
    return V3{t.x * w + c.x + p.x, t.y * w + c.y + p.y, t.z * w + c.z + p.z};
}
PCV_GHD V3 iso_apply(const double* iso7, V3 p) {
    const V3 r = quat_rot(iso7, p);
    return V3{r.x + iso7[0], r.y + iso7[1], r.z + iso7[2]};
}

// What the device needs per location.
struct QueryGeom {
    int32_t kind;
    int32_t naxes;       // cached separating axes (<= 26); 0 for AllPoints
    double axes[26][3];
    double corners[8][3];
    double aabb_min[3], aabb_max[3];
    double clip_from_query[16];
    double obb_from_query[7];
    double half_extent[3];
};

inline V3 unit(V3 v) {
    const double n = std::sqrt(v3dot(v, v));
    return V3{v.x / n, v.y / n, v.z / n};
}

struct PolyIntersector {
    V3 corners[8];
    V3 edges[12];
    int nedges = 0;
    V3 normals[6];
    int nnormals = 0;
};

inline PolyIntersector frustum_intersector(const double* query_from_clip) {
    PolyIntersector r;
    int i = 0;
    for (int sx = -1; sx <= 1; sx += 2)
        for (int sy = -1; sy <= 1; sy += 2)
            for (int sz = -1; sz <= 1; sz += 2) r.corners[i++] = mat4_transform_point(query_from_clip, V3{(double)sx, (double)sy, (double)sz});
    const V3* c = r.corners;
    r.edges[0] = unit(v3sub(c[4], c[0]));
    r.edges[1] = unit(v3sub(c[2], c[0]));
    r.edges[2] = unit(v3sub(c[1], c[0]));
    r.edges[3] = unit(v3sub(c[3], c[2]));
    r.edges[4] = unit(v3sub(c[5], c[4]));
    r.edges[5] = unit(v3sub(c[7], c[6]));
    r.nedges = 6;
    r.normals[0] = unit(v3cross(r.edges[0], r.edges[1]));
    r.normals[1] = unit(v3cross(r.edges[0], r.edges[2]));
    r.normals[2] = unit(v3cross(r.edges[0], r.edges[3]));
    r.normals[3] = unit(v3cross(r.edges[1], r.edges[2]));
    r.normals[4] = unit(v3cross(r.edges[1], r.edges[4]));
    r.nnormals = 5;
    return r;
}

inline PolyIntersector obb_intersector(const double* query_from_obb, const double* h) {
    PolyIntersector r;
    for (int i = 0; i < 8; ++i) {
        const V3 l{(i & 1) ? h[0] : -h[0], (i & 2) ? h[1] : -h[1], (i & 4) ? h[2] : -h[2]};
        r.corners[i] = iso_apply(query_from_obb, l);
    }
    r.edges[0] = unit(quat_rot(query_from_obb, V3{1, 0, 0}));
    r.edges[1] = unit(quat_rot(query_from_obb, V3{0, 1, 0}));
    r.edges[2] = unit(quat_rot(query_from_obb, V3{0, 0, 1}));
    r.nedges = 3;
    for (int i = 0; i < 3; ++i) r.normals[i] = r.edges[i];
    r.nnormals = 3;
    return r;
}

// Intersector::cache_separating_axes_for_aabb: [own normals, x,y,z, normalize(edge_i x unit_j) if finite], then dedup.
inline void cache_axes_for_aabb(const PolyIntersector& p, QueryGeom& g) {
    V3 all[6 + 3 + 36];
    int n = 0;
    const V3 units[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < p.nnormals; ++i) all[n++] = p.normals[i];
    for (int j = 0; j < 3; ++j) all[n++] = units[j];
    for (int i = 0; i < p.nedges; ++i)
        for (int j = 0; j < 3; ++j) {
            const V3 c = unit(v3cross(p.edges[i], units[j]));
            if (std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z)) all[n++] = c;
        }
    g.naxes = 0;
    for (int i = 0; i < n; ++i) {
        bool dupe = false;
        for (int k = 0; k < g.naxes && !dupe; ++k) {
            const V3 o{g.axes[k][0], g.axes[k][1], g.axes[k][2]};
            const V3 dm = v3sub(all[i], o), dp = v3add(all[i], o);
            dupe = std::fmin(v3dot(dm, dm), v3dot(dp, dp)) < std::numeric_limits<double>::epsilon();
        }
        if (!dupe) {
            g.axes[g.naxes][0] = all[i].x;
            g.axes[g.naxes][1] = all[i].y;
            g.axes[g.naxes][2] = all[i].z;
            ++g.naxes;
        }
    }
    for (int i = 0; i < 8; ++i) {
        g.corners[i][0] = p.corners[i].x;
        g.corners[i][1] = p.corners[i].y;
        g.corners[i][2] = p.corners[i].z;
    }
}

inline QueryGeom make_query_geom(const pcv_location& loc) {
    QueryGeom g;
    std::memset(&g, 0, sizeof g);
    g.kind = loc.kind;
    for (int a = 0; a < 3; ++a) {
        g.aabb_min[a] = std::fmin(loc.aabb_min[a], loc.aabb_max[a]);
        g.aabb_max[a] = std::fmax(loc.aabb_min[a], loc.aabb_max[a]);
        g.half_extent[a] = loc.half_extent[a];
    }
    std::memcpy(g.clip_from_query, loc.clip_from_query, sizeof g.clip_from_query);
    std::memcpy(g.obb_from_query, loc.obb_from_query, sizeof g.obb_from_query);
    if (loc.kind == PCV_LOC_AABB) {
        g.naxes = 3;
        for (int j = 0; j < 3; ++j) g.axes[j][j] = 1.0;
        for (int i = 0; i < 8; ++i) {  // aabb.rs:114-125: x fastest
            g.corners[i][0] = (i & 1) ? g.aabb_max[0] : g.aabb_min[0];
            g.corners[i][1] = (i & 2) ? g.aabb_max[1] : g.aabb_min[1];
            g.corners[i][2] = (i & 4) ? g.aabb_max[2] : g.aabb_min[2];
        }
    } else if (loc.kind == PCV_LOC_FRUSTUM) {
        cache_axes_for_aabb(frustum_intersector(loc.query_from_clip), g);
    } else if (loc.kind == PCV_LOC_OBB) {
        cache_axes_for_aabb(obb_intersector(loc.query_from_obb, loc.half_extent), g);
    }
    return g;
}

// 4x4 inverse by cofactors (the formula nalgebra's try_inverse uses for 4x4); false if det == 0.
inline bool mat4_try_inverse(const double* m, double* out) {
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
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0) return false;
    const double inv_det = 1.0 / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * inv_det;
    return true;
}

}  // namespace pcv
