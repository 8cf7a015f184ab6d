// ply_host.hpp — host side of the PLY input path: header -> record layout (pcv_ply_info), and the streaming loader that
// feeds raw vertex records to the GPU through a pinned staging ring.
//
// Header rules follow the reference reader (src/read_write/ply.rs:126-229 parse_header, :327-450 from_file):
//   "ply" / format <ascii|binary_little_endian|binary_big_endian> 1.0 / element <name> <count> / property <type> <name> /
//   property list .. (ignored) / comment offset: x y z / comment .. (ignored) / end_header; anything else is an error.
//   Only binary_little_endian is readable; the vertex element must exist, is read from the end of the header, and must
//   carry x, y, z.  `a`/`alpha` is skipped as one byte; uchar red/green/blue (r/g/b) form the colour; a float
//   `intensity` is the intensity attribute; properties of type int8/(u)int16/(u)int32 that are not coordinates are
//   skipped with their size; float/double/uchar properties with other names are parsed past (the octree build keeps
//   colour and intensity only, generation.rs:300).  A property name ending in a digit is rejected as in the reference.
//   64-bit integer types are rejected (the reference advances 4 bytes for them, ply.rs:267-272, which cannot be meant).
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pcv.h"
#include "build_host.hpp"

namespace pcv {

inline int ply_type_of(const std::string& s) {
    static const struct {
        const char *a, *b;
        int t;
    } names[] = {{"float", "float32", PCV_PLY_F32}, {"double", "float64", PCV_PLY_F64}, {"char", "int8", PCV_PLY_I8}, {"uchar", "uint8", PCV_PLY_U8},
                 {"short", "int16", PCV_PLY_I16}, {"ushort", "uint16", PCV_PLY_U16}, {"int", "int32", PCV_PLY_I32}, {"uint", "uint32", PCV_PLY_U32},
                 {"longlong", "int64", PCV_PLY_I64}, {"ulonglong", "uint64", PCV_PLY_U64}};
    for (const auto& n : names)
        if (s == n.a || s == n.b) return n.t;
    return -1;
}
inline uint32_t ply_type_size(int t) {
    static const uint32_t sz[] = {1, 1, 2, 2, 4, 4, 8, 8, 4, 8};
    return sz[t];
}

struct PlyFile {  // RAII file descriptor
    int fd = -1;
    explicit PlyFile(const char* path) { fd = ::open(path, O_RDONLY); }
    ~PlyFile() {
        if (fd >= 0) ::close(fd);
    }
    PlyFile(const PlyFile&) = delete;
    PlyFile& operator=(const PlyFile&) = delete;
};

// Reads the header text (everything up to and including the "end_header" line) into `text`.
inline void ply_header_text(int fd, std::string& text) {
    char buf[4096];
    off_t pos = 0;
    for (;;) {
        const ssize_t got = ::pread(fd, buf, sizeof buf, pos);
        if (got < 0) throw BuildError(PCV_ERR_IO, std::string("read failed: ") + std::strerror(errno));
        if (got == 0) return;  // the parser reports the missing end_header
        text.append(buf, (size_t)got);
        pos += got;
        // stop once a complete end_header line is inside the text
        size_t at = 0;
        while ((at = text.find("end_header", at)) != std::string::npos) {
            // only the first token of a line terminates the header (ply.rs:139-146 splits lines into tokens): a
            // "comment ... end_header" line does not
            size_t b = at;
            while (b > 0 && (text[b - 1] == ' ' || text[b - 1] == '\t' || text[b - 1] == '\r')) --b;
            const bool line_start = b == 0 || text[b - 1] == '\n';
            const size_t nl = text.find('\n', at);
            if (line_start && nl != std::string::npos) return;
            at += 10;
        }
        if (text.size() > (1u << 24)) throw BuildError(PCV_ERR_INVALID, "PLY header larger than 16 MiB");
    }
}

inline pcv_ply_info ply_parse(const char* path) {
    PlyFile f(path);
    if (f.fd < 0) throw BuildError(PCV_ERR_IO, "Could not open input file.");
    std::string text;
    ply_header_text(f.fd, text);

    pcv_ply_info info{};
    size_t pos = 0;
    int lineno = 0, format = -1;
    bool done = false, in_vertex = false, have_vertex = false, in_element = false;
    bool seen[3] = {false, false, false}, col[3] = {false, false, false};
    uint32_t off = 0;
    while (!done) {
        // one line, including its terminator (read_line semantics: a last line without '\n' still counts)
        const size_t nl = text.find('\n', pos);
        const size_t end = nl == std::string::npos ? text.size() : nl + 1;
        const std::string line = text.substr(pos, end - pos);
        pos = end;
        info.header_bytes += line.size();
        ++lineno;
        std::vector<std::string> w;
        for (size_t i = 0; i < line.size();) {
            while (i < line.size() && std::isspace((unsigned char)line[i])) ++i;
            size_t j = i;
            while (j < line.size() && !std::isspace((unsigned char)line[j])) ++j;
            if (j > i) w.push_back(line.substr(i, j - i));
            i = j;
        }
        if (lineno == 1) {
            if (w.size() != 1 || w[0] != "ply") throw BuildError(PCV_ERR_INVALID, "Not a PLY file");
            continue;
        }
        const auto bad_line = [&]() { return BuildError(PCV_ERR_INVALID, "Invalid line: " + line); };
        if (w.empty()) throw bad_line();
        if (w[0] == "format" && w.size() == 3) {
            if (w[2] != "1.0") throw BuildError(PCV_ERR_INVALID, "Invalid version: " + w[2]);
            if (w[1] == "binary_little_endian") format = 0;
            else if (w[1] == "binary_big_endian") format = 1;
            else if (w[1] == "ascii") format = 2;
            else throw BuildError(PCV_ERR_INVALID, "Invalid format: " + w[1]);
        } else if (w[0] == "element" && w.size() == 3) {
            char* e = nullptr;
            errno = 0;
            const long long cnt = std::strtoll(w[2].c_str(), &e, 10);
            if (errno || !e || *e || e == w[2].c_str()) throw BuildError(PCV_ERR_INVALID, "Invalid count: " + w[2]);
            in_element = true;
            in_vertex = !have_vertex && w[1] == "vertex";  // header["vertex"]: the first element of that name
            if (in_vertex) {
                have_vertex = true;
                if (cnt < 0) throw BuildError(PCV_ERR_INVALID, "Invalid count: " + w[2]);
                info.num_points = (uint64_t)cnt;
            }
        } else if (w[0] == "property") {
            if (!in_element) throw BuildError(PCV_ERR_INVALID, "property outside of element: " + line);
            if (w.size() == 5 && w[1] == "list") continue;
            if (w.size() != 3) throw bad_line();
            const int t = ply_type_of(w[1]);
            if (t < 0) throw BuildError(PCV_ERR_INVALID, "Invalid data type: " + w[1]);
            if (!in_vertex) continue;
            const std::string& name = w[2];
            const bool wide_int = t == PCV_PLY_I64 || t == PCV_PLY_U64;
            uint32_t size = ply_type_size(t);
            const int axis = name == "x" ? 0 : name == "y" ? 1 : name == "z" ? 2 : -1;
            if (axis >= 0) {
                if (wide_int) throw BuildError(PCV_ERR_UNSUPPORTED, "64-bit integer coordinates are not supported");
                seen[axis] = true;
                info.type_xyz[axis] = t;
                info.off_xyz[axis] = off;
            } else if (name == "a" || name == "alpha") {
                size = 1;
            } else {
                if (std::isdigit((unsigned char)name.back()))
                    throw BuildError(PCV_ERR_UNSUPPORTED, "Multidimensional attributes other than position and color are currently unsupported.");
                if (wide_int) throw BuildError(PCV_ERR_UNSUPPORTED, "64-bit integer properties are not supported");
                const int ch = (name == "r" || name == "red") ? 0 : (name == "g" || name == "green") ? 1 : (name == "b" || name == "blue") ? 2 : -1;
                if (ch >= 0) {
                    if (t == PCV_PLY_U8) {
                        col[ch] = true;
                        info.off_rgb[ch] = off;
                    } else if (t == PCV_PLY_F32 || t == PCV_PLY_F64) {
                        throw BuildError(PCV_ERR_INVALID, "colour channels must be uchar");
                    }
                } else if (name == "intensity" && t == PCV_PLY_F32) {
                    info.has_intensity = 1;
                    info.off_intensity = off;
                }
            }
            off += size;
        } else if (w[0] == "end_header") {
            done = true;
        } else if (w[0] == "comment") {
            if (w.size() == 5 && w[1] == "offset:") {
                for (int a = 0; a < 3; ++a) {
                    char* e = nullptr;
                    info.offset[a] = std::strtod(w[2 + a].c_str(), &e);
                    if (!e || *e || e == w[2 + a].c_str()) throw BuildError(PCV_ERR_INVALID, "Invalid offset: " + w[2 + a]);
                }
            }
        } else {
            throw bad_line();
        }
        if (!done && pos >= text.size()) throw BuildError(PCV_ERR_INVALID, "Invalid line: ");  // end of file before end_header
    }
    if (format < 0) throw BuildError(PCV_ERR_INVALID, "No format specified");
    if (!have_vertex) throw BuildError(PCV_ERR_INVALID, "Header does not have element 'vertex'");
    if (format != 0) throw BuildError(PCV_ERR_UNSUPPORTED, "Unsupported PLY format: only binary_little_endian");
    if (!seen[0] || !seen[1] || !seen[2]) throw BuildError(PCV_ERR_INVALID, "PLY must contain properties 'x', 'y', 'z' for 'vertex'.");
    if (col[0] && !(col[1] && col[2])) throw BuildError(PCV_ERR_INVALID, "colour needs red, green and blue");
    info.has_color = col[0] ? 1 : 0;
    info.record_bytes = off;
    return info;
}

inline void ply_validate(const pcv_ply_info& i) {
    // one 256-record tile must fit the unpack kernel's shared memory (256 * record_bytes + colour staging <= 200 KB)
    if (i.record_bytes == 0) throw BuildError(PCV_ERR_INVALID, "PLY record size is zero");
    if (i.record_bytes > 768) throw BuildError(PCV_ERR_UNSUPPORTED, "PLY vertex records larger than 768 bytes are not supported");
    for (int a = 0; a < 3; ++a) {
        const int t = i.type_xyz[a];
        if (t < 0 || t > PCV_PLY_F64 || t == PCV_PLY_I64 || t == PCV_PLY_U64) throw BuildError(PCV_ERR_INVALID, "bad coordinate type");
        if (i.off_xyz[a] + ply_type_size(t) > i.record_bytes) throw BuildError(PCV_ERR_INVALID, "coordinate outside the record");
        if (i.has_color && i.off_rgb[a] + 1 > i.record_bytes) throw BuildError(PCV_ERR_INVALID, "colour outside the record");
    }
    if (i.has_intensity && i.off_intensity + 4 > i.record_bytes) throw BuildError(PCV_ERR_INVALID, "intensity outside the record");
}

}  // namespace pcv
