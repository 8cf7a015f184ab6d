// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
//
// The S2 side of the reference: S2Splitter::write / get_meta (src/read_write/s2.rs:52-125,165-173),
// CellID::from_point (src/math/mod.rs:119-131), CellUnion as a PointCulling (src/geometry/s2_cell_union.rs:27-31) and
// S2Cells::nodes_in_location for AllPoints and S2Cells (src/s2_cells/mod.rs:157-168,233-241).
//
// THIRD-PARTY, UN-VENDORED: the cell arithmetic is the `s2` crate's (0.0.10 in the reference's Cargo.lock; a port of
// golang/geo), which is absent from /root/reference and cannot be fetched.  It is restated from the published S2 algorithm in
// the libraries' own formulation (PointFromCoords, face / validFaceXYZToUV, the quadratic uvToST, stToIJ, the 1024-entry
// lookupPos table built by initLookupCell, cellIDFromFaceIJ, Parent, RangeMin / RangeMax, ContainsCellID, IntersectsCellID,
// Normalize, ToToken).  PARITY UNPINNED: the reference holds no S2 golden vector (its tests compare S2 results with its own
// octree results), so this restatement is checked against the structural invariants S2 publishes instead
// (tests/test_s2.py): the six face cells and their tokens, the leaf cell of (1, 0, 0), parents containing children, the
// Hilbert curve visiting edge-adjacent cells at every level and across the face boundaries, round trips through the cell centre.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace orc {
namespace s2 {

constexpr int kMaxLevel = 30, kLookupBits = 4, kSwapMask = 1, kInvertMask = 2;
constexpr int kPosToIJ[4][4] = {{0, 1, 3, 2}, {0, 2, 3, 1}, {3, 2, 0, 1}, {3, 1, 0, 2}};
constexpr int kPosToOrientation[4] = {kSwapMask, 0, 0, kInvertMask | kSwapMask};

struct Lookup {
    int pos[1 << (2 * kLookupBits + 2)];
    int ij[1 << (2 * kLookupBits + 2)];
    Lookup() {
        init(0, 0, 0, 0, 0, 0);
        init(0, 0, 0, kSwapMask, 0, kSwapMask);
        init(0, 0, 0, kInvertMask, 0, kInvertMask);
        init(0, 0, 0, kSwapMask | kInvertMask, 0, kSwapMask | kInvertMask);
    }
    void init(int level, int i, int j, int orig_orientation, int p, int orientation) {  // initLookupCell
        if (level == kLookupBits) {
            const int ijv = (i << kLookupBits) + j;
            pos[(ijv << 2) + orig_orientation] = (p << 2) + orientation;
            ij[(p << 2) + orig_orientation] = (ijv << 2) + orientation;
            return;
        }
        level++;
        i <<= 1;
        j <<= 1;
        p <<= 2;
        const int* r = kPosToIJ[orientation];
        for (int k = 0; k < 4; ++k) init(level, i + (r[k] >> 1), j + (r[k] & 1), orig_orientation, p + k, orientation ^ kPosToOrientation[k]);
    }
};
inline const Lookup& lookup() {
    static const Lookup l;
    return l;
}

struct V3 {
    double x, y, z;
};
inline V3 point_from_coords(double x, double y, double z) {  // PointFromCoords / Vector::normalize
    if (x == 0. && y == 0. && z == 0.) return V3{-0.0099994664350250197, 0.0025924542609324121, 0.99994664350250195};
    const double norm = std::sqrt(x * x + y * y + z * z);
    const double m = 1.0 / norm;
    return V3{x * m, y * m, z * m};
}
inline int largest_component(const V3& v) {
    const double ax = std::fabs(v.x), ay = std::fabs(v.y), az = std::fabs(v.z);
    if (ax > ay) {
        if (ax > az) return 0;
        return 2;
    }
    if (ay > az) return 1;
    return 2;
}
inline int face_of(const V3& r) {
    int f = largest_component(r);
    if ((f == 0 && r.x < 0) || (f == 1 && r.y < 0) || (f == 2 && r.z < 0)) f += 3;
    return f;
}
inline void valid_face_xyz_to_uv(int face, const V3& r, double& u, double& v) {
    switch (face) {
        case 0: u = r.y / r.x, v = r.z / r.x; return;
        case 1: u = -r.x / r.y, v = r.z / r.y; return;
        case 2: u = -r.x / r.z, v = -r.y / r.z; return;
        case 3: u = r.z / r.x, v = r.y / r.x; return;
        case 4: u = r.z / r.y, v = -r.x / r.y; return;
        default: u = -r.y / r.z, v = -r.x / r.z; return;
    }
}
inline double uv_to_st(double u) {
    if (u >= 0) return 0.5 * std::sqrt(1 + 3 * u);
    return 1 - 0.5 * std::sqrt(1 - 3 * u);
}
inline double st_to_uv(double s) {
    if (s >= 0.5) return (1 / 3.) * (4 * s * s - 1);
    return (1 / 3.) * (1 - 4 * (1 - s) * (1 - s));
}
inline int st_to_ij(double s) {
    const double f = std::floor((double)(1 << kMaxLevel) * s);
    int64_t v = f != f ? 0 : (f >= 2147483647.0 ? 2147483647 : (f <= -2147483648.0 ? -2147483648ll : (int64_t)f));
    return (int)std::min<int64_t>(std::max<int64_t>(v, 0), (1 << kMaxLevel) - 1);
}
inline uint64_t from_face_ij(int f, int i, int j) {  // cellIDFromFaceIJ
    uint64_t n = (uint64_t)f << 60;
    int bits = f & kSwapMask;
    const Lookup& L = lookup();
    for (int k = 7; k >= 0; --k) {
        const int mask = (1 << kLookupBits) - 1;
        bits += ((i >> (k * kLookupBits)) & mask) << (kLookupBits + 2);
        bits += ((j >> (k * kLookupBits)) & mask) << 2;
        bits = L.pos[bits];
        n |= (uint64_t)(bits >> 2) << (k * 2 * kLookupBits);
        bits &= (kSwapMask | kInvertMask);
    }
    return n * 2 + 1;
}
inline uint64_t cell_id_from_point(double x, double y, double z) {
    const V3 p = point_from_coords(x, y, z);
    const int f = face_of(p);
    double u, v;
    valid_face_xyz_to_uv(f, p, u, v);
    return from_face_ij(f, st_to_ij(uv_to_st(u)), st_to_ij(uv_to_st(v)));
}
inline uint64_t lsb(uint64_t id) { return id & (0 - id); }
inline uint64_t lsb_for_level(int level) { return 1ull << (2 * (kMaxLevel - level)); }
inline uint64_t parent(uint64_t id, int level) {
    const uint64_t l = lsb_for_level(level);
    return (id & (0 - l)) | l;
}
inline int level_of(uint64_t id) { return kMaxLevel - (__builtin_ctzll(id) >> 1); }
inline uint64_t range_min(uint64_t id) { return id - (lsb(id) - 1); }
inline uint64_t range_max(uint64_t id) { return id + (lsb(id) - 1); }
inline bool contains(uint64_t a, uint64_t b) { return range_min(a) <= b && b <= range_max(a); }
inline uint64_t next(uint64_t id) { return id + (lsb(id) << 1); }  // CellID::next
inline std::string to_token(uint64_t id) {
    if (id == 0) return "X";
    char buf[17];
    snprintf(buf, sizeof buf, "%016llx", (unsigned long long)id);
    std::string s(buf);
    while (!s.empty() && s.back() == '0') s.pop_back();
    return s;
}
// faceIJOrientation: (face, i, j) of a cell id (the inverse walk; used by the structural tests)
inline void face_ij(uint64_t id, int& f, int& i, int& j) {
    f = (int)(id >> 61);
    int bits = f & kSwapMask;
    i = j = 0;
    const Lookup& L = lookup();
    for (int k = 7; k >= 0; --k) {
        const int nbits = k == 7 ? kMaxLevel - 7 * kLookupBits : kLookupBits;
        bits += (int)((id >> (k * 2 * kLookupBits + 1)) & ((1 << (2 * nbits)) - 1)) << 2;
        bits = L.ij[bits];
        i += (bits >> (kLookupBits + 2)) << (k * kLookupBits);
        j += ((bits >> 2) & ((1 << kLookupBits) - 1)) << (k * kLookupBits);
        bits &= (kSwapMask | kInvertMask);
    }
}
// The centre of a cell as a direction (face_uv_to_xyz of the centre (s, t)); not normalised.
inline V3 cell_centre_raw(uint64_t id) {
    int f, i, j;
    face_ij(id, f, i, j);
    const int lvl = level_of(id);
    const int size = 1 << (kMaxLevel - lvl);
    const int ci = (i & -size) * 2 + size, cj = (j & -size) * 2 + size;  // centre in units of 2^-31
    const double u = st_to_uv((double)ci / 2147483648.0), v = st_to_uv((double)cj / 2147483648.0);
    switch (f) {
        case 0: return V3{1, u, v};
        case 1: return V3{-u, 1, v};
        case 2: return V3{-u, -v, 1};
        case 3: return V3{-1, -v, -u};
        case 4: return V3{v, -1, -u};
        default: return V3{v, u, -1};
    }
}

// CellUnion over sorted, normalised ids
inline void normalize(std::vector<uint64_t>& cu) {
    std::sort(cu.begin(), cu.end());
    std::vector<uint64_t> out;
    auto siblings = [](uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
        if ((a ^ b ^ c) != d) return false;
        uint64_t mask = lsb(d) << 1;
        mask = ~(mask + (mask << 1));
        const uint64_t idm = d & mask;
        return (a & mask) == idm && (b & mask) == idm && (c & mask) == idm && level_of(d) != 0;
    };
    for (uint64_t ci : cu) {
        if (!out.empty() && contains(out.back(), ci)) continue;
        while (!out.empty() && contains(ci, out.back())) out.pop_back();
        while (out.size() >= 3 && siblings(out[out.size() - 3], out[out.size() - 2], out[out.size() - 1], ci)) {
            out.resize(out.size() - 3);
            ci = parent(ci, level_of(ci) - 1);
        }
        out.push_back(ci);
    }
    cu.swap(out);
}
inline bool union_contains(const std::vector<uint64_t>& cu, uint64_t id) {  // ContainsCellID
    const size_t i = std::lower_bound(cu.begin(), cu.end(), id) - cu.begin();
    if (i < cu.size() && range_min(cu[i]) <= id) return true;
    return i != 0 && range_max(cu[i - 1]) >= id;
}
inline bool union_intersects(const std::vector<uint64_t>& cu, uint64_t id) {  // IntersectsCellID
    const size_t i = std::lower_bound(cu.begin(), cu.end(), id) - cu.begin();
    if (i < cu.size() && range_min(cu[i]) <= range_max(id)) return true;
    return i != 0 && range_max(cu[i - 1]) >= range_min(id);
}

// S2Splitter::write over one stream of points + get_meta: cells in id order, every cell's points in input order.
struct SplitResult {
    bool ok = true;
    uint64_t bad_index = 0;  // first point that "is not a valid ECEF point"
    double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    std::map<uint64_t, std::vector<uint64_t>> cells;  // cell id -> source indices
};
inline SplitResult split(const double* x, const double* y, const double* z, uint64_t stride, uint64_t n, int split_level) {
    SplitResult r;
    for (uint64_t k = 0; k < n; ++k) {
        const double px = x[k * stride], py = y[k * stride], pz = z[k * stride];
        const double radius = std::sqrt(px * px + py * py + pz * pz);  // pos.coords.norm()
        if (radius > 6384400.0 || radius < 6352800.0) {
            r.ok = false;
            r.bad_index = k;
            return r;
        }
        const double p[3] = {px, py, pz};
        for (int a = 0; a < 3; ++a) {
            r.bmin[a] = k == 0 ? p[a] : std::fmin(r.bmin[a], p[a]);
            r.bmax[a] = k == 0 ? p[a] : std::fmax(r.bmax[a], p[a]);
        }
        r.cells[parent(cell_id_from_point(px, py, pz), split_level)].push_back(k);
    }
    return r;
}

}  // namespace s2
}  // namespace orc
