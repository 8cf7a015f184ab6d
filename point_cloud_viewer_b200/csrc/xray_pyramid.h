// xray_pyramid.h — per-element logic of the X-ray pipeline beyond the leaf tile, shared by the CUDA kernels
// (xray_pyramid.cuh) and the sequential test backend (tests/cpu_backend): every function here is `PCV_HD`, so the arithmetic
// that the `-m "not gpu"` tests check against the oracle is the arithmetic the kernels run.
//
// Reference (file:line relative to the reference checkout):
//   binned columns       xray/src/generation.rs:129-157 (bins), :210-363 (strategies)
//   build_parent         xray/src/generation.rs:410-451 (2 x 2 mosaic, child 1 top left, 0 bottom left, 3 top right, 2 bottom right)
//   build_node           xray/src/generation.rs:722-759 (mosaic -> image::imageops::resize(.., Lanczos3))
//   assign_background    xray/src/generation.rs:695-720 (alpha < 128 -> background colour)
//   quadtree ids / rects quadtree/src/lib.rs:57-141,143-230
// `image` 0.23.10 is an un-vendored dependency; its resize is restated from the published 0.23 source (see
// oracle/oracle_xray_pyramid.hpp for the statement and what is and is not pinned).
// Compile with FMA contraction off (nvcc -fmad=false / g++ -ffp-contract=off): `t += v * w` is a multiply then an add.
#pragma once
#include <stdint.h>

#include "chain.h"
#include "lod_order.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace pcv {

// ---- resampling (image 0.23 imageops::sample) ---------------------------------------------------------
// One output sample = a window [left, left + count) of input samples, f32 weights, their sum (accumulated in window
// order).  The table is computed on the HOST (f32 `sin` must be glibc's sinf, as Rust's f32::sin is on linux-gnu) and read
// by the kernels; for the 2:1 reduction of a parent tile the window is at most 13 samples.
struct ResampleTaps {      // device view
    const uint32_t* left;  // [out]
    const uint32_t* first; // [out] offset of the sample's weights in w
    const uint32_t* count; // [out]
    const float* sum;      // [out]
    const float* w;
};

PCV_HD uint8_t img_f32_to_u8(float v) {  // NumCast::from(FloatNearest(clamp(v, 0, 255))): f32::round = half away from zero
    const float c = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
    return (uint8_t)roundf(c);
}

// The 2N x 2N mosaic of build_parent, never materialised: pixel (x, y) comes from one of the children or is the background.
struct MosaicSrc {
    const uint8_t* child[4];  // RGBA, cs x cs, or null
    uint32_t cs;              // child edge in pixels
    uint32_t bg;              // background colour r | g << 8 | b << 16 | a << 24
};
PCV_HD uint32_t load_rgba(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
PCV_HD uint32_t mosaic_pixel(const MosaicSrc& m, uint32_t x, uint32_t y) {
    const bool right = x >= m.cs, bottom = y >= m.cs;
    const int id = right ? (bottom ? 2 : 3) : (bottom ? 0 : 1);
    const uint8_t* c = m.child[id];
    if (!c) return m.bg;
    const uint32_t lx = right ? x - m.cs : x, ly = bottom ? y - m.cs : y;
    return load_rgba(c + ((size_t)ly * m.cs + lx) * 4);
}

PCV_HD uint32_t resample_finish(const float acc[4], float sum) {
    return (uint32_t)img_f32_to_u8(acc[0] / sum) | ((uint32_t)img_f32_to_u8(acc[1] / sum) << 8) | ((uint32_t)img_f32_to_u8(acc[2] / sum) << 16) |
           ((uint32_t)img_f32_to_u8(acc[3] / sum) << 24);
}
// vertical_sample of one pixel: column x of the mosaic, output row oy.
PCV_HD uint32_t resample_v_pixel(const MosaicSrc& m, const ResampleTaps& t, uint32_t x, uint32_t oy) {
    const uint32_t left = t.left[oy], n = t.count[oy];
    const float* w = t.w + t.first[oy];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t p = mosaic_pixel(m, x, left + i);
        const float wi = w[i];
        acc[0] += (float)(p & 255u) * wi;
        acc[1] += (float)((p >> 8) & 255u) * wi;
        acc[2] += (float)((p >> 16) & 255u) * wi;
        acc[3] += (float)(p >> 24) * wi;
    }
    return resample_finish(acc, t.sum[oy]);
}
// horizontal_sample of one pixel: row `row` (in_w RGBA pixels as u32), output column ox.
PCV_HD uint32_t resample_h_pixel(const uint32_t* row, const ResampleTaps& t, uint32_t ox) {
    const uint32_t left = t.left[ox], n = t.count[ox];
    const float* w = t.w + t.first[ox];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t p = row[left + i];
        const float wi = w[i];
        acc[0] += (float)(p & 255u) * wi;
        acc[1] += (float)((p >> 8) & 255u) * wi;
        acc[2] += (float)((p >> 16) & 255u) * wi;
        acc[3] += (float)(p >> 24) * wi;
    }
    return resample_finish(acc, t.sum[ox]);
}

PCV_HD uint32_t background_pixel(uint32_t p, uint32_t bg) { return (p >> 24) < 128u ? bg : p; }  // generation.rs:711

// ---- binned columns -------------------------------------------------------------------------------------
PCV_HD int64_t rust_f64_as_i64(double v) {  // Rust `as i64`: truncate, saturate, NaN -> 0
    if (!(v == v)) return 0;
    if (v >= 9223372036854775808.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)v;
}
PCV_HD int64_t xray_bin_of(float attr, double bin_size) { return rust_f64_as_i64((double)attr / bin_size); }  // generation.rs:143-145

#if defined(__CUDA_ARCH__)
#define PCV_CAS64(p, c, v) atomicCAS((unsigned long long*)(p), (unsigned long long)(c), (unsigned long long)(v))
#define PCV_ADDF(p, v) atomicAdd((p), (v))
#define PCV_ADDU(p, v) atomicAdd((p), (v))
#define PCV_SETI(p, v) atomicExch((p), (v))
#else
inline uint64_t pcv_seq_cas64(uint64_t* p, uint64_t c, uint64_t v) {
    const uint64_t old = *p;
    if (old == c) *p = v;
    return old;
}
#define PCV_CAS64(p, c, v) pcv_seq_cas64((uint64_t*)(p), (uint64_t)(c), (uint64_t)(v))
#define PCV_ADDF(p, v) (*(p) += (v))
#define PCV_ADDU(p, v) (*(p) += (v))
#define PCV_SETI(p, v) (*(p) = (v))
#endif

// Hash aggregation of (pixel, bin) -> (sums, count).  Two open-addressing tables, keys claimed with one 64-bit CAS each
// (no locks, no spinning on another thread's progress):
//   bins:    i64 bin -> dense id = its slot.  The empty marker is the bit pattern of i64::MIN; the bin i64::MIN itself (only
//            reachable by saturation) has the dedicated id `bin_cap`.
//   columns: (pixel << 32 | bin id) -> slot; the empty marker is ~0 (a pixel index is < w * h <= 2^32 - 1).
// Both tables are sized so that they cannot fill up through valid use (columns: > 2 x the points of the tile); a full bin
// table (more distinct bins than `bin_cap`) raises `*err`.
constexpr uint64_t kBinEmpty = 0x8000000000000000ull;
constexpr uint64_t kColEmpty = ~0ull;
struct BinnedTables {
    uint64_t* bin_keys;  // [bin_cap], power of two
    uint32_t bin_cap;
    uint64_t* col_keys;  // [col_cap]
    uint64_t col_cap;
    float* col_sum;      // [col_cap * ncomp]
    uint32_t* col_count; // [col_cap]
    int ncomp;           // 3 (colour) or 1 (intensity)
    int* err;
};
PCV_HD uint32_t binned_bin_id(const BinnedTables& t, int64_t bin) {  // returns ~0u on a full table
    if ((uint64_t)bin == kBinEmpty) return t.bin_cap;
    const uint32_t mask = t.bin_cap - 1u;
    uint32_t h = (uint32_t)lod_mix64((uint64_t)bin) & mask;
    for (uint32_t probe = 0; probe < t.bin_cap; ++probe) {
        uint64_t prev = *(volatile const uint64_t*)&t.bin_keys[h];  // a claimed key never changes: most look-ups need no atomic
        if (prev == kBinEmpty) prev = PCV_CAS64(&t.bin_keys[h], kBinEmpty, (uint64_t)bin);
        if (prev == kBinEmpty || prev == (uint64_t)bin) return h;
        h = (h + 1u) & mask;
    }
    return 0xFFFFFFFFu;
}
PCV_HD void binned_insert(const BinnedTables& t, uint32_t pixel, int64_t bin, const float* value) {
    const uint32_t id = binned_bin_id(t, bin);
    if (id == 0xFFFFFFFFu) {
        PCV_SETI(t.err, 1);
        return;
    }
    const uint64_t key = ((uint64_t)pixel << 32) | (uint64_t)id;
    uint64_t h = lod_mix64(key) % t.col_cap;
    for (uint64_t probe = 0; probe < t.col_cap; ++probe) {
        uint64_t prev = *(volatile const uint64_t*)&t.col_keys[h];
        if (prev == kColEmpty) prev = PCV_CAS64(&t.col_keys[h], kColEmpty, key);
        if (prev == kColEmpty || prev == key) {
            for (int k = 0; k < t.ncomp; ++k) PCV_ADDF(&t.col_sum[h * (uint64_t)t.ncomp + k], value[k]);
            PCV_ADDU(&t.col_count[h], 1u);
            return;
        }
        h = h + 1 == t.col_cap ? 0 : h + 1;
    }
    PCV_SETI(t.err, 2);
}
// One occupied column slot -> its pixel: the bin's mean joins the pixel's sum over bins (generation.rs:276-283, :339-346).
// pix_sum has `stride` floats per pixel (4 for colour, 1 for intensity: the layout k_xray_resolve_attr reads).
PCV_HD void binned_reduce_slot(const BinnedTables& t, uint64_t slot, float* pix_sum, int stride, uint32_t* pix_bins) {
    const uint64_t key = t.col_keys[slot];
    if (key == kColEmpty) return;
    const uint32_t pixel = (uint32_t)(key >> 32);
    const float n = (float)t.col_count[slot];
    for (int k = 0; k < t.ncomp; ++k) PCV_ADDF(&pix_sum[(size_t)pixel * stride + k], t.col_sum[slot * (uint64_t)t.ncomp + k] / n);
    PCV_ADDU(&pix_bins[pixel], 1u);
}

// ---- quadtree ids and rectangles (quadtree/src/lib.rs) ---------------------------------------------------
struct QuadId {
    uint8_t level;
    uint64_t index;
};
inline bool operator<(const QuadId& a, const QuadId& b) { return a.level != b.level ? a.level < b.level : a.index < b.index; }
inline QuadId quad_child(const QuadId& p, int k) { return QuadId{(uint8_t)(p.level + 1), (p.index << 2) + (uint64_t)k}; }  // lib.rs:163-168
inline QuadId quad_parent(const QuadId& c) { return QuadId{(uint8_t)(c.level - 1), c.index >> 2}; }                         // lib.rs:178-186
struct QuadRect {
    double min_x, min_y, edge;
};
inline QuadRect quad_child_rect(const QuadRect& r, int k) {  // Node::get_child, lib.rs:84-101: bit 0 -> +y, bit 1 -> +x
    const double half = r.edge / 2.;
    QuadRect c{r.min_x, r.min_y, half};
    if (k & 1) c.min_y += half;
    if (k & 2) c.min_x += half;
    return c;
}
inline QuadRect quad_rect_of(const QuadId& id, const QuadRect& root) {  // Node::from_node_id_and_root_bounding_rect, lib.rs:62-82
    QuadRect r = root;
    for (int l = (int)id.level - 1; l >= 0; --l) r = quad_child_rect(r, (int)((id.index >> (2 * l)) & 3));
    return r;
}

// Host: the taps of every output sample of one axis (sample.rs horizontal_sample / vertical_sample).
struct ResampleTable {
    std::vector<uint32_t> left, first, count;
    std::vector<float> sum, w;
};
inline float img_sinc(float t) {
    const float a = t * 3.14159274101257324f;  // f32::consts::PI
    return t == 0.0f ? 1.0f : sinf(a) / a;
}
inline float img_lanczos3(float x) { return fabsf(x) < 3.0f ? img_sinc(x) * img_sinc(x / 3.0f) : 0.0f; }
inline ResampleTable make_lanczos3_table(uint32_t in_size, uint32_t out_size) {
    ResampleTable t;
    const float ratio = (float)in_size / (float)out_size;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float src_support = 3.0f * sratio;
    for (uint32_t o = 0; o < out_size; ++o) {
        float input = ((float)o + 0.5f) * ratio;
        int64_t left = (int64_t)floorf(input - src_support);
        left = std::min<int64_t>(std::max<int64_t>(left, 0), (int64_t)in_size - 1);
        int64_t right = (int64_t)ceilf(input + src_support);
        right = std::min<int64_t>(std::max<int64_t>(right, left + 1), (int64_t)in_size);
        input = input - 0.5f;
        t.left.push_back((uint32_t)left);
        t.first.push_back((uint32_t)t.w.size());
        t.count.push_back((uint32_t)(right - left));
        float sum = 0.f;
        for (int64_t i = left; i < right; ++i) {
            const float w = img_lanczos3(((float)i - input) / sratio);
            t.w.push_back(w);
            sum += w;
        }
        t.sum.push_back(sum);
    }
    return t;
}
// find_quadtree_bounding_rect_and_levels, xray/src/generation.rs:515-533.  Returns false where the reference's loop would
// not end or its u8 level counter would overflow (tile size not a positive finite number, more than 255 doublings).
inline bool quadtree_rect_and_levels(const double bmin[3], const double bmax[3], uint32_t tile_size_px, double pixel_size_m, QuadRect& rect, uint8_t& levels) {
    const double tile_size_m = (double)tile_size_px * pixel_size_m;
    if (!(tile_size_m > 0.0) || !(tile_size_m < INFINITY)) return false;
    int l = 0;
    double cur = tile_size_m;
    const double dx = bmax[0] - bmin[0], dy = bmax[1] - bmin[1];
    while (cur < dx || cur < dy) {
        cur *= 2.;
        if (++l > 255) return false;
    }
    levels = (uint8_t)l;
    rect = QuadRect{bmin[0], bmin[1], cur};
    return true;
}

}  // namespace pcv
