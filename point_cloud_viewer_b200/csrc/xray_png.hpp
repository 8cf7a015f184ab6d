// xray_png.hpp — host side of the X-ray quadtree's on-disk form (SURVEY 8 f3 "PNG encode", which stays on the host):
// `<node id>.png` per tile (xray/src/utils.rs:33-37 get_image_path, IMAGE_FILE_EXTENSION = "png"; RGBA, 8 bit) and the
// quadtree's meta file (xray/src/lib.rs:88-139 Meta::to_disk / to_proto, xray_proto_rust/src/proto.proto:22-56; file name =
// the root node's id with the "r" prefix replaced by "meta", + ".pb": utils.rs:7-11).  The PNG stream is a plain
// non-interlaced RGBA image with filter type 0 on every scanline, deflated by zlib: the pixels are what the reference's
// `image.save` stores, the bytes of the file are not (another encoder).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "disk_io.hpp"

namespace pcv {

inline void png_put32(std::string& s, uint32_t v) {
    const char b[4] = {(char)(v >> 24), (char)(v >> 16), (char)(v >> 8), (char)v};
    s.append(b, 4);
}
inline void png_chunk(std::string& out, const char type[4], const std::string& data) {
    png_put32(out, (uint32_t)data.size());
    std::string body(type, 4);
    body += data;
    out += body;
    png_put32(out, (uint32_t)crc32(0L, (const Bytef*)body.data(), (uInt)body.size()));
}
inline bool encode_png_rgba(const uint8_t* rgba, uint32_t w, uint32_t h, std::string& out, int level = 3) {
    if (w == 0 || h == 0 || (uint64_t)w * h * 4 + h >= 0xFFFFFFFFull) return false;
    std::vector<uint8_t> raw((size_t)h * ((size_t)w * 4 + 1));
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t* row = &raw[(size_t)y * ((size_t)w * 4 + 1)];
        row[0] = 0;  // filter type None
        memcpy(row + 1, rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf cap = compressBound((uLong)raw.size());
    std::string z(cap, '\0');
    if (compress2((Bytef*)&z[0], &cap, raw.data(), (uLong)raw.size(), level) != Z_OK) return false;
    z.resize(cap);
    out.assign("\x89PNG\r\n\x1a\n", 8);
    std::string ihdr;
    png_put32(ihdr, w);
    png_put32(ihdr, h);
    ihdr += std::string("\x08\x06\x00\x00\x00", 5);  // 8 bit, colour type 6 (RGBA), deflate, adaptive filtering, no interlace
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", std::string());
    return true;
}

// quadtree NodeId Display (quadtree/src/lib.rs:216-233): "r" + one base-4 digit per level, most significant first
inline std::string quad_node_name(uint8_t level, uint64_t index) {
    std::string s = "r";
    for (int l = (int)level - 1; l >= 0; --l) s.push_back((char)('0' + ((index >> (2 * l)) & 3)));
    return s;
}

struct XrayMetaData {
    double min_x = 0, min_y = 0, edge = 0;
    uint32_t deepest_level = 0, tile_size = 0;
    std::vector<std::pair<uint32_t, uint64_t>> nodes;  // (level, index)
};
// xray Meta::to_proto (lib.rs:117-139): version = CURRENT_VERSION = 3
inline std::string encode_xray_meta(const XrayMetaData& m) {
    std::string v2;  // Vector2d { 1: x, 2: y }
    pb::put_double(v2, 1, m.min_x);
    pb::put_double(v2, 2, m.min_y);
    std::string rect;  // Rect { 3: min, 4: edge_length }
    pb::put_bytes(rect, 3, v2);
    pb::put_double(rect, 4, m.edge);
    std::string meta;  // Meta { 1: version, 2: bounding_rect, 3: deepest_level, 4: tile_size, 5: repeated NodeId { 1: level, 2: index } }
    pb::put_uint(meta, 1, 3);
    pb::put_bytes(meta, 2, rect);
    pb::put_uint(meta, 3, m.deepest_level);
    pb::put_uint(meta, 4, m.tile_size);
    for (const auto& n : m.nodes) {
        std::string id;
        pb::put_uint(id, 1, n.first);
        pb::put_uint(id, 2, n.second);
        pb::put_bytes(meta, 5, id);
    }
    return meta;
}

}  // namespace pcv
