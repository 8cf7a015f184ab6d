// s2.h — the S2 cell arithmetic the S2-cell point cloud needs, shared by the CUDA kernels (s2.cuh) and the
// sequential test backend.  Everything is integer or IEEE-754 binary64 (+, *, /, sqrt, floor): no libm, so host and device
// agree bit for bit.
//
// Reference call sites (file:line relative to the reference checkout):
//   CellID::from_point(p) = CellID::from(Point::from_coords(x, y, z))   src/math/mod.rs:119-131
//   S2Splitter::write: radius check, bounding box, from_point(p).parent(split_level), per-cell batches
//                                                                      src/read_write/s2.rs:14-17,59-125
//   CellUnion as PointCulling: contains_cellid(from_point(p))          src/geometry/s2_cell_union.rs:27-31
//   S2Cells::nodes_in_location (AllPoints, S2Cells)                     src/s2_cells/mod.rs:157-168,233-241
// THIRD-PARTY, UN-VENDORED: the arithmetic itself lives in the `s2` crate (0.0.10 in Cargo.lock, a port of golang/geo's s2
// package), which is not in /root/reference.  It is restated here from the published S2 algorithm (the same in the C++, Go and
// Rust libraries): point -> unit vector -> face and (u, v) by the largest component -> quadratic (s, t) -> (i, j) in
// [0, 2^30) -> position along the Hilbert curve of the face.  This header walks the curve level by level with the 4 x 4
// tables; the oracle (oracle/oracle_s2.hpp) uses the libraries' 1024-entry look-up tables - two formulations of one curve.
#pragma once
#include <stdint.h>

#include "chain.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

namespace pcv {

constexpr int kS2MaxLevel = 30;
constexpr double kEarthRadiusMinM = 6352800.0;  // src/math/mod.rs:30-35
constexpr double kEarthRadiusMaxM = 6384400.0;

// The orientation change after each curve position (the libraries' posToOrientation; swapMask = 1, invertMask = 2).
PCV_HD int s2_pos_to_orientation(int pos) { return pos == 0 ? 1 : pos == 3 ? 3 : 0; }

// uv -> st (quadratic projection) and st -> ij
PCV_HD double s2_uv_to_st(double u) { return u >= 0.0 ? 0.5 * sqrt(1.0 + 3.0 * u) : 1.0 - 0.5 * sqrt(1.0 - 3.0 * u); }
PCV_HD int s2_st_to_ij(double s) {
    const double f = floor(1073741824.0 * s);  // maxSize = 2^30
    // clamp(int(f), 0, maxSize - 1); a NaN becomes 0 (Rust `as i32`; unreachable for points that pass the radius check)
    if (!(f >= 0.0)) return 0;
    return f >= 1073741823.0 ? 1073741823 : (int)f;
}

// Point::from_coords + xyz_to_face_uv: the unit vector, its face and (u, v).
PCV_HD int s2_face_uv(double x, double y, double z, double& u, double& v) {
    if (!(x == 0.0 && y == 0.0 && z == 0.0)) {  // Vector::normalize: v * (1 / |v|)
        const double inv = 1.0 / sqrt(x * x + y * y + z * z);
        x = x * inv, y = y * inv, z = z * inv;
    } else {  // Point::origin(): (-0.0099994664, 0.0025924542, 0.9999466403), already unit length in the libraries' eyes
        x = -0.0099994664350250197, y = 0.0025924542609324121, z = 0.99994664350250195;
    }
    const double ax = fabs(x), ay = fabs(y), az = fabs(z);
    int axis;  // largest_component: ties go to the later axis
    if (ax > ay)
        axis = ax > az ? 0 : 2;
    else
        axis = ay > az ? 1 : 2;
    const double c = axis == 0 ? x : axis == 1 ? y : z;
    const int face = c < 0.0 ? axis + 3 : axis;
    switch (face) {
        case 0: u = y / x, v = z / x; break;
        case 1: u = -x / y, v = z / y; break;
        case 2: u = -x / z, v = -y / z; break;
        case 3: u = z / x, v = y / x; break;
        case 4: u = z / y, v = -x / y; break;
        default: u = -y / z, v = -x / z; break;
    }
    return face;
}

// cellIDFromFaceIJ: the leaf cell (level 30) of (face, i, j), walking the Hilbert curve one level at a time.
PCV_HD uint64_t s2_from_face_ij(int face, int i, int j) {
    uint64_t n = (uint64_t)face << 60;
    int orientation = face & 1;
    for (int k = kS2MaxLevel - 1; k >= 0; --k) {
        const int ij = (((i >> k) & 1) << 1) | ((j >> k) & 1);
        // position of the sub-cell (i bit, j bit) along the curve in the current orientation
        int pos;
        switch (orientation) {
            case 0: pos = ij == 0 ? 0 : ij == 1 ? 1 : ij == 3 ? 2 : 3; break;  // posToIJ[0] = {0, 1, 3, 2}
            case 1: pos = ij == 0 ? 0 : ij == 2 ? 1 : ij == 3 ? 2 : 3; break;  // posToIJ[1] = {0, 2, 3, 1}
            case 2: pos = ij == 3 ? 0 : ij == 2 ? 1 : ij == 0 ? 2 : 3; break;  // posToIJ[2] = {3, 2, 0, 1}
            default: pos = ij == 3 ? 0 : ij == 1 ? 1 : ij == 0 ? 2 : 3; break; // posToIJ[3] = {3, 1, 0, 2}
        }
        n |= (uint64_t)pos << (2 * k);
        orientation ^= s2_pos_to_orientation(pos);
    }
    return n * 2 + 1;
}

PCV_HD uint64_t s2_cell_id_from_point(double x, double y, double z) {
    double u, v;
    const int face = s2_face_uv(x, y, z, u, v);
    return s2_from_face_ij(face, s2_st_to_ij(s2_uv_to_st(u)), s2_st_to_ij(s2_uv_to_st(v)));
}

PCV_HD uint64_t s2_lsb_for_level(int level) { return 1ull << (2 * (kS2MaxLevel - level)); }
PCV_HD uint64_t s2_lsb(uint64_t id) { return id & (0 - id); }
PCV_HD uint64_t s2_parent(uint64_t id, int level) {  // CellID::parent(level)
    const uint64_t lsb = s2_lsb_for_level(level);
    return (id & (0 - lsb)) | lsb;
}
PCV_HD uint64_t s2_range_min(uint64_t id) { return id - (s2_lsb(id) - 1); }
PCV_HD uint64_t s2_range_max(uint64_t id) { return id + (s2_lsb(id) - 1); }
PCV_HD int s2_level(uint64_t id) {  // 30 - (trailing zeros) / 2
    int tz = 0;
    while (tz < 64 && !((id >> tz) & 1)) ++tz;
    return kS2MaxLevel - tz / 2;
}
PCV_HD bool s2_is_valid(uint64_t id) { return (id >> 61) < 6 && (s2_lsb(id) & 0x1555555555555555ull) != 0; }

// CellUnion::contains_cellid / intersects_cellid over a NORMALISED union (sorted, no cell contains another):
// binary search for the first cell >= id, then the two range checks of the libraries.
PCV_HD uint32_t s2_lower_bound(const uint64_t* cells, uint32_t n, uint64_t id) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cells[mid] < id)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
PCV_HD bool s2_union_contains(const uint64_t* cells, uint32_t n, uint64_t id) {
    const uint32_t i = s2_lower_bound(cells, n, id);
    if (i < n && s2_range_min(cells[i]) <= id) return true;
    return i != 0 && s2_range_max(cells[i - 1]) >= id;
}
PCV_HD bool s2_union_intersects(const uint64_t* cells, uint32_t n, uint64_t id) {
    const uint32_t i = s2_lower_bound(cells, n, id);
    if (i < n && s2_range_min(cells[i]) <= s2_range_max(id)) return true;
    return i != 0 && s2_range_max(cells[i - 1]) >= s2_range_min(id);
}

// The S2Splitter's validity rule (read_write/s2.rs:64-71): |p| outside [EARTH_RADIUS_MIN_M, EARTH_RADIUS_MAX_M] is an error.
// nalgebra's norm(): sqrt of the sum of squares in x, y, z order.
PCV_HD bool s2_valid_ecef(double x, double y, double z) {
    const double r = sqrt(x * x + y * y + z * z);
    return !(r > kEarthRadiusMaxM || r < kEarthRadiusMinM);
}

// ---- host helpers ------------------------------------------------------------------------------------------------
// CellID::to_token: the id in hex without its trailing zero digits; "X" for 0.
inline std::string s2_to_token(uint64_t id) {
    if (id == 0) return "X";
    char buf[17];
    snprintf(buf, sizeof buf, "%016llx", (unsigned long long)id);
    std::string s(buf);
    while (!s.empty() && s.back() == '0') s.pop_back();
    return s;
}
inline bool s2_from_token(const std::string& t, uint64_t& id) {
    if (t == "X") {
        id = 0;
        return true;
    }
    if (t.empty() || t.size() > 16) return false;
    uint64_t v = 0;
    for (char ch : t) {
        int d = ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : -1;
        if (d < 0) return false;
        v = (v << 4) | (uint64_t)d;
    }
    id = v << (4 * (16 - t.size()));
    return true;
}
// CellUnion::normalize: sort, drop cells contained in an earlier one, replace four sibling cells by their parent.
inline void s2_normalize(std::vector<uint64_t>& cells) {
    std::sort(cells.begin(), cells.end());
    std::vector<uint64_t> out;
    for (uint64_t id : cells) {
        if (!out.empty() && s2_range_min(out.back()) <= id && id <= s2_range_max(out.back())) continue;  // contained in the previous cell
        while (!out.empty() && s2_range_min(id) <= out.back() && out.back() <= s2_range_max(id)) out.pop_back();  // contains previous cells
        while (out.size() >= 3 && s2_level(id) != 0) {  // the last three cells + id: the four children of one parent?  (areSiblings)
            const size_t m = out.size();
            const uint64_t a = out[m - 3], b = out[m - 2], c = out[m - 1];
            if ((a ^ b ^ c) != id) break;  // the libraries' fast reject: the XOR of four siblings is zero
            const uint64_t two = s2_lsb(id) << 1, mask = ~(two + (two << 1));  // everything above the two bits that number the children
            const uint64_t idm = id & mask;
            if ((a & mask) != idm || (b & mask) != idm || (c & mask) != idm) break;
            out.resize(m - 3);
            id = s2_parent(id, s2_level(id) - 1);
        }
        out.push_back(id);
    }
    cells.swap(out);
}

}  // namespace pcv
