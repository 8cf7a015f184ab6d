// TEST-ONLY: sequential drivers of csrc/s2.h (the PCV_HD functions the S2 kernels call).  NOT part of the shipped library.
#include <cstring>
#include <vector>

#include "../../point_cloud_viewer_b200/csrc/s2.h"

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
}
