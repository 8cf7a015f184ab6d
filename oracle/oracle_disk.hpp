// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
//
// meta.pb (proto3 wire format of point_viewer_proto_rust/src/proto.proto:58-149, version 13) and the
// on-disk node layout of src/data_provider/on_disk.rs:17-33 / src/lib.rs:74-80, hand-encoded.
#pragma once
#include <cstdio>
#include <fstream>

#include "oracle_core.hpp"

namespace orc {

struct PbWriter {
    std::string s;
    void varint(uint64_t v) {
        while (v >= 0x80) {
            s.push_back((char)(v | 0x80));
            v >>= 7;
        }
        s.push_back((char)v);
    }
    void tag(int field, int wt) { varint(((uint64_t)field << 3) | (uint64_t)wt); }
    void f64(int field, double d) {  // proto3: default (0.0, all-zero bits) omitted
        uint64_t b;
        std::memcpy(&b, &d, 8);
        if (b == 0) return;
        tag(field, 1);
        for (int i = 0; i < 8; ++i) s.push_back((char)(b >> (8 * i)));
    }
    void u64(int field, uint64_t v) {
        if (v == 0) return;
        tag(field, 0);
        varint(v);
    }
    void msg(int field, const std::string& m) {
        tag(field, 2);
        varint(m.size());
        s += m;
    }
};

inline std::string vec3d_pb(Vec3 v) {  // proto.proto:32-36
    PbWriter w;
    w.f64(1, v.x);
    w.f64(2, v.y);
    w.f64(3, v.z);
    return w.s;
}

inline std::string meta_pb(const Octree& o) {  // octree/mod.rs:87-99, node.rs:260-270
    PbWriter cuboid;
    cuboid.msg(3, vec3d_pb(o.bbox.mins));
    cuboid.msg(4, vec3d_pb(o.bbox.maxs));
    PbWriter om;
    om.f64(2, o.resolution);
    for (auto& kv : o.nodes) {
        PbWriter nid;
        nid.u64(3, kv.first.high());
        nid.u64(4, kv.first.low());
        PbWriter node;
        node.u64(2, (uint64_t)kv.second.enc);
        node.u64(3, (uint64_t)kv.second.num_points);
        node.msg(4, nid.s);  // rust-protobuf always writes a set message field, even when empty
        om.msg(3, node.s);
    }
    PbWriter meta;
    meta.u64(1, 13);  // CURRENT_VERSION, src/lib.rs:48
    meta.msg(4, cuboid.s);
    meta.msg(6, om.s);
    return meta.s;
}

struct PbReader {
    const uint8_t* p;
    const uint8_t* e;
    bool ok = true;
    uint64_t varint() {
        uint64_t v = 0;
        int sh = 0;
        while (p < e) {
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << sh;
            if (!(b & 0x80)) return v;
            sh += 7;
        }
        ok = false;
        return v;
    }
    double f64() {
        uint64_t b = 0;
        if (e - p < 8) {
            ok = false;
            return 0;
        }
        for (int i = 0; i < 8; ++i) b |= (uint64_t)p[i] << (8 * i);
        p += 8;
        double d;
        std::memcpy(&d, &b, 8);
        return d;
    }
    PbReader sub() {
        uint64_t n = varint();
        PbReader r{p, p + n};
        if ((uint64_t)(e - p) < n) {
            ok = false;
            r.e = e;
        }
        p += n;
        return r;
    }
    void skip(int wt) {
        if (wt == 0)
            varint();
        else if (wt == 1)
            p += 8;
        else if (wt == 2)
            sub();
        else if (wt == 5)
            p += 4;
        else
            ok = false;
    }
};

inline Vec3 parse_vec3d(PbReader r) {
    Vec3 v{0, 0, 0};
    while (r.p < r.e && r.ok) {
        uint64_t t = r.varint();
        int f = (int)(t >> 3), wt = (int)(t & 7);
        if (wt == 1 && f >= 1 && f <= 3) {
            double d = r.f64();
            (f == 1 ? v.x : f == 2 ? v.y : v.z) = d;
        } else
            r.skip(wt);
    }
    return v;
}

inline bool parse_meta_pb(const std::string& buf, Octree& o) {  // octree/mod.rs:156-215 (version 13 only)
    PbReader r{(const uint8_t*)buf.data(), (const uint8_t*)buf.data() + buf.size()};
    int version = 0;
    Vec3 bmin{0, 0, 0}, bmax{0, 0, 0};
    struct N {
        NodeId id;
        int64_t n;
        int enc;
    };
    std::vector<N> nodes;
    while (r.p < r.e && r.ok) {
        uint64_t t = r.varint();
        int f = (int)(t >> 3), wt = (int)(t & 7);
        if (f == 1 && wt == 0)
            version = (int)r.varint();
        else if (f == 4 && wt == 2) {
            PbReader c = r.sub();
            while (c.p < c.e && c.ok) {
                uint64_t t2 = c.varint();
                int f2 = (int)(t2 >> 3), w2 = (int)(t2 & 7);
                if (f2 == 3 && w2 == 2)
                    bmin = parse_vec3d(c.sub());
                else if (f2 == 4 && w2 == 2)
                    bmax = parse_vec3d(c.sub());
                else
                    c.skip(w2);
            }
        } else if (f == 6 && wt == 2) {
            PbReader om = r.sub();
            while (om.p < om.e && om.ok) {
                uint64_t t2 = om.varint();
                int f2 = (int)(t2 >> 3), w2 = (int)(t2 & 7);
                if (f2 == 2 && w2 == 1)
                    o.resolution = om.f64();
                else if (f2 == 3 && w2 == 2) {
                    PbReader nd = om.sub();
                    N n{NodeId(), 0, 0};
                    while (nd.p < nd.e && nd.ok) {
                        uint64_t t3 = nd.varint();
                        int f3 = (int)(t3 >> 3), w3 = (int)(t3 & 7);
                        if (f3 == 2 && w3 == 0)
                            n.enc = (int)nd.varint();
                        else if (f3 == 3 && w3 == 0)
                            n.n = (int64_t)nd.varint();
                        else if (f3 == 4 && w3 == 2) {
                            PbReader idr = nd.sub();
                            uint64_t hi = 0, lo = 0;
                            while (idr.p < idr.e && idr.ok) {
                                uint64_t t4 = idr.varint();
                                int f4 = (int)(t4 >> 3), w4 = (int)(t4 & 7);
                                if (f4 == 3 && w4 == 0)
                                    hi = idr.varint();
                                else if (f4 == 4 && w4 == 0)
                                    lo = idr.varint();
                                else
                                    idr.skip(w4);
                            }
                            n.id = NodeId::from_high_low(hi, lo);
                        } else
                            nd.skip(w3);
                    }
                    nodes.push_back(n);
                } else
                    om.skip(w2);
            }
        } else
            r.skip(wt);
    }
    if (!r.ok || version != 13) return false;
    o.bbox = Aabb::make(bmin, bmax);
    Cube root = o.root_cube();
    for (auto& n : nodes) {
        NodeMeta m;
        m.num_points = n.n;
        m.enc = (Enc)n.enc;
        m.cube = find_bounding_cube(n.id, root);
        o.nodes[n.id] = m;
    }
    return true;
}

inline bool write_file(const std::string& path, const void* data, size_t n) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    size_t w = n ? std::fwrite(data, 1, n, f) : 0;
    std::fclose(f);
    return w == n;
}
inline bool read_file(const std::string& path, std::string& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t r = n ? std::fread(&out[0], 1, (size_t)n, f) : 0;
    std::fclose(f);
    return r == (size_t)n;
}

inline bool write_dir(const Octree& o, const std::string& dir) {
    for (auto& kv : o.files) {
        const NodeFile& f = kv.second;
        if (f.rgb.empty()) continue;
        std::string stem = dir + "/" + kv.first.to_string();
        if (!write_file(stem + ".xyz", f.xyz.data(), f.xyz.size())) return false;
        if (!write_file(stem + ".rgb", f.rgb.data(), f.rgb.size())) return false;
        if (o.with_intensity && !write_file(stem + ".intensity", f.intensity.data(), f.intensity.size() * 4)) return false;
    }
    std::string m = meta_pb(o);
    return write_file(dir + "/meta.pb", m.data(), m.size());
}

inline bool load_dir(const std::string& dir, Octree& o) {
    std::string m;
    if (!read_file(dir + "/meta.pb", m)) return false;
    if (!parse_meta_pb(m, o)) return false;
    o.with_intensity = false;
    for (auto& kv : o.nodes) {
        if (kv.second.num_points == 0) continue;
        NodeFile f;
        f.enc = kv.second.enc;
        f.cube = kv.second.cube;
        std::string stem = dir + "/" + kv.first.to_string(), b;
        if (!read_file(stem + ".xyz", b)) return false;
        f.xyz.assign(b.begin(), b.end());
        if (!read_file(stem + ".rgb", b)) return false;
        f.rgb.assign(b.begin(), b.end());
        if (read_file(stem + ".intensity", b)) {
            o.with_intensity = true;
            f.intensity.resize(b.size() / 4);
            std::memcpy(f.intensity.data(), b.data(), f.intensity.size() * 4);
        }
        f.src.assign((size_t)f.num_points(), 0);
        o.files[kv.first] = std::move(f);
    }
    return true;
}

}  // namespace orc
