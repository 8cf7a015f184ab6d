// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the point_cloud_viewer hot path (octree build, node codec, SAT culling,
// LOD node selection, filtered point queries, X-ray tile accumulation).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may link or call
// anything in this directory; the product library (point_cloud_viewer_b200/csrc) never does.
//
// The reference is Rust (no rustc/cargo in this image), so it cannot be compiled here; this file
// restates the algorithm literally, function by function, citing the reference file:line it follows
// (paths relative to the reference checkout).  Third-party arithmetic that is not vendored in the
// reference tree (nalgebra 0.22.0, simba 0.2.1, num 0.3.0, std BinaryHeap) is restated from the
// published behaviour of those crates; see DESIGN.md "Oracle pinning" for what the reference's own
// tests pin (node-id algebra, SAT relations, OBB axis counts, frustum-contains case, the 100 001
// point octree) and what they leave unpinned (encode rounding mode, ulp-level op order).
//
// Must be compiled with -ffp-contract=off: the reference never contracts a*b+c except for the two
// explicit mul_add calls in decode (src/read_write/codec.rs:130,138).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// NodeId  (src/octree/node.rs:52-173)
// ---------------------------------------------------------------------------------------------
struct NodeId {
    u128 v = 0;
    static NodeId from_level_index(uint8_t level, u128 index) {  // node.rs:108-111
        NodeId n;
        n.v = ((u128)level << 120) | index;
        return n;
    }
    uint8_t level() const { return (uint8_t)(v >> 120); }  // node.rs:147-149
    u128 index() const { return v & ((((u128)1) << 120) - 1); }  // node.rs:152-154
    NodeId child(unsigned k) const {  // node.rs:120-125
        return from_level_index(level() + 1, (index() << 3) + k);
    }
    bool has_parent() const { return level() != 0; }
    NodeId parent() const {  // node.rs:136-144
        return from_level_index(level() - 1, index() >> 3);
    }
    int child_index() const { return level() == 0 ? -1 : (int)(index() & 7); }  // node.rs:128-133
    uint64_t high() const { return (uint64_t)(v >> 64); }  // node.rs:101-106
    uint64_t low() const { return (uint64_t)v; }
    static NodeId from_high_low(uint64_t hi, uint64_t lo) {
        NodeId n;
        n.v = ((u128)hi << 64) | lo;
        return n;
    }
    bool operator<(const NodeId& o) const { return v < o.v; }
    bool operator==(const NodeId& o) const { return v == o.v; }
    std::string to_string() const {  // node.rs:73-86: 'r' + zero padded octal of width level
        std::string s = "r";
        int L = level();
        for (int i = L - 1; i >= 0; --i) s.push_back((char)('0' + (int)((index() >> (3 * i)) & 7)));
        return s;
    }
    static NodeId from_string(const std::string& s) {  // node.rs:58-71
        uint8_t level = (uint8_t)(s.size() - 1);
        u128 idx = 0;
        for (size_t i = 1; i < s.size(); ++i) idx = (idx << 3) | (u128)(s[i] - '0');
        return from_level_index(level, idx);
    }
};

struct Vec3 {
    double x, y, z;
};
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }

// ---------------------------------------------------------------------------------------------
// Aabb / Cube  (src/geometry/aabb.rs)
// ---------------------------------------------------------------------------------------------
struct Aabb {
    Vec3 mins, maxs;
    static Aabb make(Vec3 a, Vec3 b) {  // aabb.rs:19-24 (inf / sup)
        Aabb r;
        r.mins = {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)};
        r.maxs = {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)};
        return r;
    }
    void grow(Vec3 p) {  // aabb.rs:41-44
        mins = {std::fmin(mins.x, p.x), std::fmin(mins.y, p.y), std::fmin(mins.z, p.z)};
        maxs = {std::fmax(maxs.x, p.x), std::fmax(maxs.y, p.y), std::fmax(maxs.z, p.z)};
    }
    bool contains(Vec3 p) const {  // aabb.rs:46-48: mins <= p (all) && p < maxs (all)
        return mins.x <= p.x && mins.y <= p.y && mins.z <= p.z && p.x < maxs.x && p.y < maxs.y &&
               p.z < maxs.z;
    }
    Vec3 diag() const { return maxs - mins; }  // aabb.rs:54-56
    void corners(Vec3 c[8]) const {  // aabb.rs:114-125
        c[0] = {mins.x, mins.y, mins.z};
        c[1] = {maxs.x, mins.y, mins.z};
        c[2] = {mins.x, maxs.y, mins.z};
        c[3] = {maxs.x, maxs.y, mins.z};
        c[4] = {mins.x, mins.y, maxs.z};
        c[5] = {maxs.x, mins.y, maxs.z};
        c[6] = {mins.x, maxs.y, maxs.z};
        c[7] = {maxs.x, maxs.y, maxs.z};
    }
};

struct Cube {
    Vec3 min;
    double edge;
    static Cube bounding(const Aabb& b) {  // aabb.rs:149-157
        double e = std::fmax(std::fmax(b.maxs.x - b.mins.x, b.maxs.y - b.mins.y), b.maxs.z - b.mins.z);
        return Cube{b.mins, e};
    }
    Vec3 max() const { return {min.x + edge, min.y + edge, min.z + edge}; }  // aabb.rs:175-181
    Vec3 center() const {  // aabb.rs:184-192
        Vec3 mx = max();
        return {(min.x + mx.x) / 2., (min.y + mx.y) / 2., (min.z + mx.z) / 2.};
    }
    Aabb to_aabb() const { return Aabb::make(min, max()); }  // aabb.rs:159-161
};

inline Cube find_bounding_cube(NodeId id, const Cube& root) {  // node.rs:157-172
    double e = root.edge;
    Vec3 m = root.min;
    for (int level = (int)id.level() - 1; level >= 0; --level) {
        e /= 2.;
        unsigned ci = (unsigned)((id.v >> (3 * level)) & 7);
        unsigned z = ci & 1, y = (ci >> 1) & 1, x = (ci >> 2) & 1;
        m.x += (double)x * e;
        m.y += (double)y * e;
        m.z += (double)z * e;
    }
    return Cube{m, e};
}

inline Cube get_child_cube(const Cube& c, unsigned k) {  // node.rs:190-211 (Node::get_child)
    double h = c.edge / 2.;
    Vec3 m = c.min;
    if (k & 1) m.z += h;
    if (k & 2) m.y += h;
    if (k & 4) m.x += h;
    return Cube{m, h};
}

inline unsigned child_index_of(const Cube& cube, Vec3 v) {  // node.rs:34-42
    Vec3 c = cube.center();
    unsigned gx = v.x > c.x, gy = v.y > c.y, gz = v.z > c.z;
    return gx << 2 | gy << 1 | gz;
}

// ---------------------------------------------------------------------------------------------
// Position codec  (src/read_write/codec.rs)
// ---------------------------------------------------------------------------------------------
enum Enc { ENC_U8 = 1, ENC_U16 = 2, ENC_F32 = 3, ENC_F64 = 4 };  // proto.proto:78-84

inline uint32_t rust_f64_as_u32(double v) {  // Rust `as u32`: truncate, saturate, NaN -> 0
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
inline int64_t rust_f64_as_i64(double v) {
    if (!(v == v)) return 0;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    if (v >= 9223372036854775808.0) return INT64_MAX;
    return (int64_t)v;
}

inline Enc position_encoding(const Cube& cube, double resolution) {  // codec.rs:31-40
    uint32_t min_bits = rust_f64_as_u32(std::log2(cube.edge / resolution)) + 1;
    if (min_bits <= 8) return ENC_U8;
    if (min_bits <= 16) return ENC_U16;
    if (min_bits <= 24) return ENC_F32;
    return ENC_F64;
}
inline int bytes_per_coordinate(Enc e) { return e == ENC_U8 ? 1 : e == ENC_U16 ? 2 : e == ENC_F32 ? 4 : 8; }

inline double num_clamp(double x, double lo, double hi) {  // num 0.3.0 clamp
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

// Encode one coordinate; result returned as the raw little-endian bits in a u64.
inline uint64_t encode_coord(double value, double min, double edge, Enc enc) {
    double t = num_clamp((value - min) / edge, 0., 1.);  // codec.rs:111,119,142-148
    switch (enc) {
        case ENC_U8: {  // codec.rs:102-113 ; simba SubsetOf<f64> for u8 == `as u8`
            double s = 255.0 * t;
            uint32_t v = rust_f64_as_u32(s);
            return v > 255u ? 255u : v;
        }
        case ENC_U16: {
            double s = 65535.0 * t;
            uint32_t v = rust_f64_as_u32(s);
            return v > 65535u ? 65535u : v;
        }
        case ENC_F32: {  // codec.rs:115-121 ; `as f32` rounds to nearest even
            float f = (float)t;
            uint32_t b;
            std::memcpy(&b, &f, 4);
            return b;
        }
        default: {
            uint64_t b;
            std::memcpy(&b, &t, 8);
            return b;
        }
    }
}

inline double decode_coord(uint64_t bits, double min, double edge, Enc enc) {
    switch (enc) {
        case ENC_U8: return std::fma((double)(uint32_t)bits / 255.0, edge, min);  // codec.rs:124-131
        case ENC_U16: return std::fma((double)(uint32_t)bits / 65535.0, edge, min);
        case ENC_F32: {  // codec.rs:134-139
            float f;
            uint32_t b = (uint32_t)bits;
            std::memcpy(&f, &b, 4);
            return std::fma((double)f, edge, min);
        }
        default: {
            double d;
            std::memcpy(&d, &bits, 8);
            return std::fma(d, edge, min);
        }
    }
}

inline void store_le(uint8_t* dst, uint64_t bits, int nbytes) {
    for (int i = 0; i < nbytes; ++i) dst[i] = (uint8_t)(bits >> (8 * i));
}
inline uint64_t load_le(const uint8_t* src, int nbytes) {
    uint64_t b = 0;
    for (int i = 0; i < nbytes; ++i) b |= (uint64_t)src[i] << (8 * i);
    return b;
}

// ---------------------------------------------------------------------------------------------
// In-memory node "files"  (layout of src/data_provider/on_disk.rs:17-33, src/lib.rs:74-80,
// src/read_write/raw.rs:374-392): .xyz = n * 3 * bpc little-endian interleaved; .rgb = n*3 u8;
// .intensity = n * f32.  `src` is an oracle-only provenance side channel (input index).
// ---------------------------------------------------------------------------------------------
struct NodeFile {
    Enc enc = ENC_U8;
    Cube cube{};
    std::vector<uint8_t> xyz;
    std::vector<uint8_t> rgb;
    std::vector<float> intensity;
    std::vector<uint64_t> src;
    size_t disk_points = 0;  // faithful build variant: the content currently lives in the node's files (oracle_build.hpp)
    int64_t num_points() const { return (int64_t)(rgb.size() / 3); }  // on_disk.rs:23-33
};

struct Point {
    Vec3 p;
    uint8_t rgb[3];
    float intensity;
    uint64_t src;
};

inline void node_append(NodeFile& f, const Point& pt, bool with_intensity) {  // raw.rs:374-392
    int bpc = bytes_per_coordinate(f.enc);
    size_t o = f.xyz.size();
    f.xyz.resize(o + 3 * bpc);
    store_le(&f.xyz[o], encode_coord(pt.p.x, f.cube.min.x, f.cube.edge, f.enc), bpc);
    store_le(&f.xyz[o + bpc], encode_coord(pt.p.y, f.cube.min.y, f.cube.edge, f.enc), bpc);
    store_le(&f.xyz[o + 2 * bpc], encode_coord(pt.p.z, f.cube.min.z, f.cube.edge, f.enc), bpc);
    f.rgb.push_back(pt.rgb[0]);
    f.rgb.push_back(pt.rgb[1]);
    f.rgb.push_back(pt.rgb[2]);
    if (with_intensity) f.intensity.push_back(pt.intensity);
    f.src.push_back(pt.src);
}

inline Point node_read(const NodeFile& f, size_t i, bool with_intensity) {  // raw.rs:127-216
    int bpc = bytes_per_coordinate(f.enc);
    const uint8_t* s = &f.xyz[i * 3 * bpc];
    Point pt;
    pt.p.x = decode_coord(load_le(s, bpc), f.cube.min.x, f.cube.edge, f.enc);
    pt.p.y = decode_coord(load_le(s + bpc, bpc), f.cube.min.y, f.cube.edge, f.enc);
    pt.p.z = decode_coord(load_le(s + 2 * bpc, bpc), f.cube.min.z, f.cube.edge, f.enc);
    pt.rgb[0] = f.rgb[3 * i];
    pt.rgb[1] = f.rgb[3 * i + 1];
    pt.rgb[2] = f.rgb[3 * i + 2];
    pt.intensity = with_intensity ? f.intensity[i] : 0.f;
    pt.src = f.src[i];
    return pt;
}

// ---------------------------------------------------------------------------------------------
// Octree (result of a build, or loaded from a directory)
// ---------------------------------------------------------------------------------------------
struct NodeMeta {
    int64_t num_points;
    Enc enc;
    Cube cube;
};

struct Octree {
    double resolution = 0;
    Aabb bbox{};
    bool with_intensity = false;
    std::map<NodeId, NodeMeta> nodes;      // meta.pb content (includes zero-point nodes)
    std::map<NodeId, NodeFile> files;      // only nodes with >= 1 point
    Cube root_cube() const { return Cube::bounding(bbox); }
};

}  // namespace orc
