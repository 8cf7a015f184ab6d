// oracle_ply.hpp — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp).  CPU restatement of the reference's PLY input path:
//   parse_header          src/read_write/ply.rs:126-229
//   PlyIterator::from_file src/read_write/ply.rs:327-450   (property readers, skipped properties, record size)
//   PlyIterator::next      src/read_write/ply.rs:522-556   (batching)
//   batch_from_readers     src/read_write/ply.rs:453-514   (position = (x,y,z) + offset; colour from r/g/b columns)
//   find_bounding_box      src/octree/generation.rs:256-270
// Pinned against the reference's own fixtures (src/test_data/*.ply, tests ply.rs:746-790) in tests/test_ply_pins.py.
//
// Quirks of the reference that are restated, not fixed:
//   * an `int8`/`char` coordinate is read as an unsigned byte (ply.rs:254: `buf[0]` for Int8);
//   * `a`/`alpha` is skipped as ONE byte whatever its declared type (ply.rs:383-385);
//   * 64-bit integer properties advance the cursor by 4 bytes while reading 8 (ply.rs:267-272).  These are not legal PLY
//     types; the restatement rejects them ("unsupported") instead of reproducing the misaligned read;
//   * the vertex element is assumed to be the first element of the body (the reader seeks to the end of the header);
//   * where the reference panics (missing vertex / x,y,z, non-little-endian, truncated body) this code reports an error.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace orc {

enum PlyType { PLY_I8, PLY_U8, PLY_I16, PLY_U16, PLY_I32, PLY_U32, PLY_I64, PLY_U64, PLY_F32, PLY_F64, PLY_INVALID };

inline PlyType ply_type_from_str(const std::string& s) {  // ply.rs:58-73
    if (s == "float" || s == "float32") return PLY_F32;
    if (s == "double" || s == "float64") return PLY_F64;
    if (s == "char" || s == "int8") return PLY_I8;
    if (s == "uchar" || s == "uint8") return PLY_U8;
    if (s == "short" || s == "int16") return PLY_I16;
    if (s == "ushort" || s == "uint16") return PLY_U16;
    if (s == "int" || s == "int32") return PLY_I32;
    if (s == "uint" || s == "uint32") return PLY_U32;
    if (s == "longlong" || s == "int64") return PLY_I64;
    if (s == "ulonglong" || s == "uint64") return PLY_U64;
    return PLY_INVALID;
}
inline int ply_type_bytes(PlyType t) {
    switch (t) {
        case PLY_I8: case PLY_U8: return 1;
        case PLY_I16: case PLY_U16: return 2;
        case PLY_I32: case PLY_U32: case PLY_F32: return 4;
        default: return 8;
    }
}

struct PlyProperty {
    std::string name;
    PlyType type;
};
struct PlyElement {
    std::string name;
    int64_t count;
    std::vector<PlyProperty> properties;
};
struct PlyHeader {
    int format = -1;  // 0 binary little endian, 1 binary big endian, 2 ascii
    std::vector<PlyElement> elements;
    double offset[3] = {0, 0, 0};
    size_t header_len = 0;
};

// One entry per vertex property, in record order.
enum PlyRole { ROLE_SKIP, ROLE_X, ROLE_Y, ROLE_Z, ROLE_R, ROLE_G, ROLE_B, ROLE_INTENSITY, ROLE_OTHER };
struct PlyField {
    std::string name;
    PlyType type;
    PlyRole role;
    uint32_t offset, bytes;  // position inside the record and bytes consumed
};
struct PlyLayout {
    PlyHeader header;
    std::vector<PlyField> fields;
    uint32_t record_bytes = 0;
    int64_t num_points = 0;
    bool has_color = false, has_intensity = false;
};

inline bool read_line(FILE* f, std::string& line) {  // BufRead::read_line: up to and including '\n'
    line.clear();
    int c;
    while ((c = fgetc(f)) != EOF) {
        line.push_back((char)c);
        if (c == '\n') break;
    }
    return !line.empty();
}
inline std::vector<std::string> split_ws(const std::string& s) {
    std::istringstream is(s);
    std::vector<std::string> out;
    std::string w;
    while (is >> w) out.push_back(w);
    return out;
}
inline bool parse_f64(const std::string& s, double& v) {
    char* end = nullptr;
    v = std::strtod(s.c_str(), &end);
    return end && *end == 0 && end != s.c_str();
}
inline bool parse_i64(const std::string& s, int64_t& v) {
    char* end = nullptr;
    v = std::strtoll(s.c_str(), &end, 10);
    return end && *end == 0 && end != s.c_str();
}

// ply.rs:126-229
inline bool ply_parse_header(FILE* f, PlyHeader& h, std::string& err) {
    std::string line;
    read_line(f, line);
    h.header_len += line.size();
    {
        auto e = split_ws(line);
        if (e.size() != 1 || e[0] != "ply") {
            err = "Not a PLY file";
            return false;
        }
    }
    bool in_element = false;
    PlyElement cur;
    for (;;) {
        read_line(f, line);
        h.header_len += line.size();
        auto e = split_ws(line);
        const std::string key = e.empty() ? "" : e[0];
        if (key == "format" && e.size() == 3) {
            if (e[2] != "1.0") {
                err = "Invalid version: " + e[2];
                return false;
            }
            if (e[1] == "ascii") h.format = 2;
            else if (e[1] == "binary_little_endian") h.format = 0;
            else if (e[1] == "binary_big_endian") h.format = 1;
            else {
                err = "Invalid format: " + e[1];
                return false;
            }
        } else if (key == "element" && e.size() == 3) {
            if (in_element) h.elements.push_back(cur);
            cur = PlyElement();
            cur.name = e[1];
            if (!parse_i64(e[2], cur.count)) {
                err = "Invalid count: " + e[2];
                return false;
            }
            in_element = true;
        } else if (key == "property") {
            if (!in_element) {
                err = "property outside of element: " + line;
                return false;
            }
            if (e.size() == 5 && e[1] == "list") continue;  // list properties are not supported: ignored
            if (e.size() == 3) {
                PlyType t = ply_type_from_str(e[1]);
                if (t == PLY_INVALID) {
                    err = "Invalid data type: " + e[1];
                    return false;
                }
                cur.properties.push_back({e[2], t});
            } else {
                err = "Invalid line: " + line;
                return false;
            }
        } else if (key == "end_header") {
            break;
        } else if (key == "comment") {
            if (e.size() == 5 && e[1] == "offset:") {
                for (int a = 0; a < 3; ++a)
                    if (!parse_f64(e[2 + a], h.offset[a])) {
                        err = "Invalid offset: " + e[2 + a];
                        return false;
                    }
            }
        } else {
            err = "Invalid line: " + line;  // includes an empty line and a premature end of file
            return false;
        }
    }
    if (in_element) h.elements.push_back(cur);
    if (h.format < 0) {
        err = "No format specified";
        return false;
    }
    return true;
}

// ply.rs:327-450
inline bool ply_open(const char* path, PlyLayout& L, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        err = "Could not open input file.";
        return false;
    }
    const bool ok = ply_parse_header(f, L.header, err);
    std::fclose(f);
    if (!ok) return false;
    const PlyElement* vertex = nullptr;
    for (const auto& e : L.header.elements)
        if (e.name == "vertex") {
            vertex = &e;
            break;
        }
    if (!vertex) {
        err = "Header does not have element 'vertex'";
        return false;
    }
    if (L.header.format != 0) {
        err = "Unsupported PLY format";
        return false;
    }
    bool sx = false, sy = false, sz = false, sr = false, sg = false, sb = false;
    uint32_t off = 0;
    for (const auto& p : vertex->properties) {
        PlyField fd{p.name, p.type, ROLE_SKIP, off, 0};
        const bool i64 = p.type == PLY_I64 || p.type == PLY_U64;
        if (p.name == "x" || p.name == "y" || p.name == "z") {
            if (i64) {
                err = "unsupported 64-bit integer coordinate";
                return false;
            }
            fd.role = p.name == "x" ? ROLE_X : p.name == "y" ? ROLE_Y : ROLE_Z;
            fd.bytes = ply_type_bytes(p.type);
            (p.name == "x" ? sx : p.name == "y" ? sy : sz) = true;
        } else if (p.name == "a" || p.name == "alpha") {
            fd.bytes = 1;  // skipped as one byte
        } else {
            if (p.name.empty() || (p.name.back() >= '0' && p.name.back() <= '9')) {
                err = "Multidimensional attributes other than position and color are currently unsupported.";
                return false;
            }
            if (i64) {
                err = "unsupported 64-bit integer property";
                return false;
            }
            fd.bytes = ply_type_bytes(p.type);
            const bool is_r = p.name == "r" || p.name == "red", is_g = p.name == "g" || p.name == "green", is_b = p.name == "b" || p.name == "blue";
            switch (p.type) {
                case PLY_U8:
                    fd.role = is_r ? ROLE_R : is_g ? ROLE_G : is_b ? ROLE_B : ROLE_OTHER;
                    break;
                case PLY_F32:
                    fd.role = p.name == "intensity" ? ROLE_INTENSITY : ROLE_OTHER;
                    break;
                case PLY_F64:
                    fd.role = ROLE_OTHER;
                    break;
                default:
                    fd.role = ROLE_SKIP;  // int8, (u)int16, (u)int32 attributes are ignored
            }
            if ((is_r || is_g || is_b) && p.type != PLY_U8) {
                // r/g/b columns of another type make batch_from_readers panic (ply.rs:463-465) unless they are skipped types
                if (p.type == PLY_F32 || p.type == PLY_F64) {
                    err = "colour channels must be uchar";
                    return false;
                }
            }
            if (fd.role == ROLE_R) sr = true;
            if (fd.role == ROLE_G) sg = true;
            if (fd.role == ROLE_B) sb = true;
            if (fd.role == ROLE_INTENSITY) L.has_intensity = true;
        }
        off += fd.bytes;
        L.fields.push_back(fd);
    }
    if (!sx || !sy || !sz) {
        err = "PLY must contain properties 'x', 'y', 'z' for 'vertex'.";
        return false;
    }
    if (sr && !(sg && sb)) {
        err = "colour needs red, green and blue";
        return false;
    }
    L.has_color = sr;
    L.record_bytes = off;
    L.num_points = vertex->count;
    return true;
}

inline double ply_as_f64(const uint8_t* p, PlyType t) {  // `$reading_fn(buf) as f64`, ply.rs:248-284
    switch (t) {
        case PLY_U8: return (double)p[0];
        case PLY_I8: return (double)p[0];  // sic: read as an unsigned byte
        case PLY_U16: { uint16_t v; std::memcpy(&v, p, 2); return (double)v; }
        case PLY_I16: { int16_t v; std::memcpy(&v, p, 2); return (double)v; }
        case PLY_U32: { uint32_t v; std::memcpy(&v, p, 4); return (double)v; }
        case PLY_I32: { int32_t v; std::memcpy(&v, p, 4); return (double)v; }
        case PLY_F32: { float v; std::memcpy(&v, p, 4); return (double)v; }
        default: { double v; std::memcpy(&v, p, 8); return v; }
    }
}

// Points [first, first + count) of the file as the PointsBatch stream would deliver them (ply.rs:453-556): SoA positions
// with the header offset added, colour interleaved r,g,b, intensity.  Batch boundaries do not influence the values.
inline bool ply_read_range(const char* path, const PlyLayout& L, uint64_t first, uint64_t count, double* x, double* y, double* z, uint8_t* rgb,
                           float* intensity, std::string& err) {
    if (first + count > (uint64_t)L.num_points) {
        err = "range exceeds the vertex count";
        return false;
    }
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        err = "Could not open input file.";
        return false;
    }
    std::fseek(f, (long)(L.header.header_len + first * L.record_bytes), SEEK_SET);
    std::vector<uint8_t> buf((size_t)L.record_bytes * 1024);  // BufReader aligned to whole points (ply.rs:440-442)
    uint64_t done = 0;
    while (done < count) {
        const uint64_t m = std::min<uint64_t>(1024, count - done);
        if (std::fread(buf.data(), L.record_bytes, m, f) != m) {
            std::fclose(f);
            err = "truncated PLY body";
            return false;
        }
        for (uint64_t i = 0; i < m; ++i) {
            const uint8_t* rec = buf.data() + i * L.record_bytes;
            const uint64_t o = done + i;
            for (const auto& fd : L.fields) {
                const uint8_t* p = rec + fd.offset;
                switch (fd.role) {
                    case ROLE_X: x[o] = ply_as_f64(p, fd.type) + L.header.offset[0]; break;
                    case ROLE_Y: y[o] = ply_as_f64(p, fd.type) + L.header.offset[1]; break;
                    case ROLE_Z: z[o] = ply_as_f64(p, fd.type) + L.header.offset[2]; break;
                    case ROLE_R: if (rgb) rgb[3 * o] = p[0]; break;
                    case ROLE_G: if (rgb) rgb[3 * o + 1] = p[0]; break;
                    case ROLE_B: if (rgb) rgb[3 * o + 2] = p[0]; break;
                    case ROLE_INTENSITY: if (intensity) std::memcpy(&intensity[o], p, 4); break;
                    default: break;
                }
            }
        }
        done += m;
    }
    std::fclose(f);
    return true;
}

// generation.rs:256-270: Aabb::new(first, first) grown by every point (component-wise min/max, aabb.rs:41-44);
// Aabb::zero for an empty file.
inline bool ply_find_bounding_box(const char* path, double mn[3], double mx[3], std::string& err) {
    PlyLayout L;
    if (!ply_open(path, L, err)) return false;
    for (int a = 0; a < 3; ++a) mn[a] = mx[a] = 0.0;
    const uint64_t n = (uint64_t)L.num_points, B = 100000;  // NUM_POINTS_PER_BATCH-sized reads
    std::vector<double> x(B), y(B), z(B);
    bool have = false;
    for (uint64_t first = 0; first < n; first += B) {
        const uint64_t m = std::min(B, n - first);
        if (!ply_read_range(path, L, first, m, x.data(), y.data(), z.data(), nullptr, nullptr, err)) return false;
        for (uint64_t i = 0; i < m; ++i) {
            const double p[3] = {x[i], y[i], z[i]};
            if (!have) {
                for (int a = 0; a < 3; ++a) mn[a] = mx[a] = p[a];
                have = true;
            }
            for (int a = 0; a < 3; ++a) {  // Aabb::grow = component-wise inf / sup, as oracle_core.hpp Aabb::grow
                mn[a] = std::fmin(mn[a], p[a]);
                mx[a] = std::fmax(mx[a], p[a]);
            }
        }
    }
    return true;
}

}  // namespace orc
