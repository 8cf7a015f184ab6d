// TEST-ONLY: sequential drivers of csrc/xray_pyramid.h - the same PCV_HD functions the CUDA kernels of xray_pyramid.cuh
// call, with the grid-stride loops replaced by plain loops (and the atomics by their sequential meaning).  Lets the
// `-m "not gpu"` tests compare the product's per-element arithmetic with the oracle on a machine without a GPU.  NOT part of
// the shipped library.
#include <cstring>
#include <vector>

#include "../../point_cloud_viewer_b200/csrc/xray_pyramid.h"
#include "../../point_cloud_viewer_b200/csrc/xray_png.hpp"

using namespace pcv;

extern "C" {

// pcv_xray_build_parent's device work: k_xray_resample_v over the virtual mosaic, then k_xray_resample_h.
int tbx_build_parent(const uint8_t* const children[4], uint32_t child_px, const uint8_t* bg4, uint32_t tile_px, uint8_t* rgba_out) {
    const ResampleTable tb = make_lanczos3_table(2 * child_px, tile_px);
    ResampleTaps t{tb.left.data(), tb.first.data(), tb.count.data(), tb.sum.data(), tb.w.data()};
    MosaicSrc m{};
    for (int k = 0; k < 4; ++k) m.child[k] = children[k];
    m.cs = child_px;
    m.bg = load_rgba(bg4);
    const uint32_t in_w = 2 * child_px;
    std::vector<uint32_t> tmp((size_t)in_w * tile_px);
    for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = resample_v_pixel(m, t, (uint32_t)(i % in_w), (uint32_t)(i / in_w));
    uint32_t* out = (uint32_t*)rgba_out;
    for (size_t i = 0; i < (size_t)tile_px * tile_px; ++i) {
        const uint32_t y = (uint32_t)(i / tile_px), ox = (uint32_t)(i % tile_px);
        out[i] = resample_h_pixel(tmp.data() + (size_t)y * in_w, t, ox);
    }
    return 0;
}

void tbx_background(uint8_t* rgba, uint64_t npix, const uint8_t* bg4) {
    uint32_t* p = (uint32_t*)rgba;
    const uint32_t bg = load_rgba(bg4);
    for (uint64_t i = 0; i < npix; ++i) p[i] = background_pixel(p[i], bg);
}

// The hash aggregation of k_xray_binned_insert + k_xray_binned_reduce over already discretised points:
// pixel[i], attr[i] (the binning attribute), value[i * ncomp ..].  pix_sum: npix * stride floats, pix_bins: npix.
int tbx_binned(uint64_t n, const uint32_t* pixel, const float* attr, const float* value, int ncomp, double bin_size, uint32_t bin_cap, uint64_t col_cap,
               uint64_t npix, int stride, float* pix_sum, uint32_t* pix_bins) {
    std::vector<uint64_t> bk(bin_cap, kBinEmpty), ck(col_cap, kColEmpty);
    std::vector<float> cs(col_cap * (size_t)ncomp, 0.f);
    std::vector<uint32_t> cc(col_cap, 0);
    int err = 0;
    BinnedTables t{bk.data(), bin_cap, ck.data(), col_cap, cs.data(), cc.data(), ncomp, &err};
    for (uint64_t i = 0; i < n; ++i) binned_insert(t, pixel[i], xray_bin_of(attr[i], bin_size), value + i * ncomp);
    std::memset(pix_sum, 0, npix * stride * sizeof(float));
    std::memset(pix_bins, 0, npix * sizeof(uint32_t));
    for (uint64_t s = 0; s < col_cap; ++s) binned_reduce_slot(t, s, pix_sum, stride, pix_bins);
    return err;
}

int64_t tbx_bin_of(float attr, double bin_size) { return xray_bin_of(attr, bin_size); }

void tbx_quad_rect_of(uint8_t level, uint64_t index, const double* root3, double* out3) {
    const QuadRect r = quad_rect_of(QuadId{level, index}, QuadRect{root3[0], root3[1], root3[2]});
    out3[0] = r.min_x, out3[1] = r.min_y, out3[2] = r.edge;
}
int tbx_rect_and_levels(const double* bmin, const double* bmax, uint32_t tile_px, double pixel_size_m, double* rect3, int* levels) {
    QuadRect r{};
    uint8_t l = 0;
    if (!quadtree_rect_and_levels(bmin, bmax, tile_px, pixel_size_m, r, l)) return -1;
    rect3[0] = r.min_x, rect3[1] = r.min_y, rect3[2] = r.edge;
    *levels = l;
    return 0;
}
// host side of the quadtree's on-disk form (csrc/xray_png.hpp)
int64_t tbx_encode_png(const uint8_t* rgba, uint32_t w, uint32_t h, uint8_t* out, uint64_t cap) {
    std::string png;
    if (!encode_png_rgba(rgba, w, h, png)) return -1;
    if (png.size() > cap) return -(int64_t)png.size();
    std::memcpy(out, png.data(), png.size());
    return (int64_t)png.size();
}
int64_t tbx_encode_xray_meta(double min_x, double min_y, double edge, uint32_t deepest, uint32_t tile, const uint32_t* levels, const uint64_t* indices, uint64_t n,
                             uint8_t* out, uint64_t cap) {
    XrayMetaData m;
    m.min_x = min_x, m.min_y = min_y, m.edge = edge, m.deepest_level = deepest, m.tile_size = tile;
    for (uint64_t k = 0; k < n; ++k) m.nodes.emplace_back(levels[k], indices[k]);
    const std::string b = encode_xray_meta(m);
    if (b.size() > cap) return -(int64_t)b.size();
    std::memcpy(out, b.data(), b.size());
    return (int64_t)b.size();
}
void tbx_node_name(uint8_t level, uint64_t index, char* buf, int cap) { snprintf(buf, cap, "%s", quad_node_name(level, index).c_str()); }
}
