// TEST-ONLY: sequential drivers of csrc/s2.h (the PCV_HD functions the S2 kernels call).  NOT part of the shipped library.
#include <cstring>
#include <vector>

#include "../../point_cloud_viewer_b200/csrc/s2.h"
#include "../../point_cloud_viewer_b200/csrc/s2_disk.hpp"

using namespace pcv;

extern "C" {
void tbs_cell_ids(uint64_t n, const double* x, const double* y, const double* z, uint64_t stride, int level, uint64_t* out, uint8_t* valid_out) {
    for (uint64_t k = 0; k < n; ++k) {
        const double px = x[k * stride], py = y[k * stride], pz = z[k * stride];
        out[k] = s2_parent(s2_cell_id_from_point(px, py, pz), level);
        if (valid_out) valid_out[k] = s2_valid_ecef(px, py, pz) ? 1 : 0;
    }
}
uint64_t tbs_from_face_ij(int f, int i, int j) { return s2_from_face_ij(f, i, j); }
uint64_t tbs_normalize(uint64_t* ids, uint64_t n) {
    std::vector<uint64_t> v(ids, ids + n);
    s2_normalize(v);
    std::memcpy(ids, v.data(), v.size() * 8);
    return v.size();
}
void tbs_union_test(const uint64_t* cu, uint32_t ncu, const uint64_t* ids, uint64_t m, uint8_t* contains_out, uint8_t* intersects_out) {
    for (uint64_t k = 0; k < m; ++k) {
        contains_out[k] = s2_union_contains(cu, ncu, ids[k]) ? 1 : 0;
        intersects_out[k] = s2_union_intersects(cu, ncu, ids[k]) ? 1 : 0;
    }
}
void tbs_token(uint64_t id, char* buf, int cap) { snprintf(buf, cap, "%s", s2_to_token(id).c_str()); }
int tbs_from_token(const char* t, uint64_t* id) { return s2_from_token(t, *id) ? 0 : -1; }
int tbs_level(uint64_t id) { return s2_level(id); }
int tbs_is_valid(uint64_t id) { return s2_is_valid(id) ? 1 : 0; }
// meta.pb of an S2 cloud (csrc/s2_disk.hpp)
int64_t tbs_encode_s2_meta(const double* bbox6, const uint64_t* ids, const uint64_t* counts, uint64_t n, int has_color, int has_intensity, uint8_t* out, uint64_t cap) {
    S2MetaData m;
    for (int a = 0; a < 3; ++a) m.bbox_min[a] = bbox6[a], m.bbox_max[a] = bbox6[3 + a];
    m.ids.assign(ids, ids + n);
    m.counts.assign(counts, counts + n);
    m.has_color = has_color != 0;
    m.has_intensity = has_intensity != 0;
    const std::string b = encode_s2_meta(m);
    if (b.size() > cap) return -(int64_t)b.size();
    std::memcpy(out, b.data(), b.size());
    return (int64_t)b.size();
}
// returns the cell count, or -1 with the message in err
int64_t tbs_decode_s2_meta(const uint8_t* buf, uint64_t len, double* bbox6, uint64_t* ids, uint64_t* counts, uint64_t cap, int* has_color, int* has_intensity,
                           int* version, char* err, int errcap) {
    S2MetaData m;
    const std::string e = decode_s2_meta(std::string((const char*)buf, len), m, *version);
    snprintf(err, errcap, "%s", e.c_str());
    if (!e.empty()) return -1;
    for (int a = 0; a < 3; ++a) bbox6[a] = m.bbox_min[a], bbox6[3 + a] = m.bbox_max[a];
    for (size_t k = 0; k < m.ids.size() && k < cap; ++k) ids[k] = m.ids[k], counts[k] = m.counts[k];
    *has_color = m.has_color;
    *has_intensity = m.has_intensity;
    return (int64_t)m.ids.size();
}
}
