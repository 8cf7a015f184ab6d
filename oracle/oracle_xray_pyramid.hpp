// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).  Never linked into, imported by or executed from the
// product; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it.
//
// The rest of the X-ray pipeline: binned columns, parent tiles, background, the quadtree driver.
// Restates, with file:line relative to the reference checkout:
//   BinnedColoringStrategy::bins                 xray/src/generation.rs:129-157
//   Intensity / PointColor strategies + binning  xray/src/generation.rs:210-363
//   build_parent                                 xray/src/generation.rs:410-451
//   find_quadtree_bounding_rect_and_levels       xray/src/generation.rs:515-533
//   get_nodes_at_level / get_bounding_box        xray/src/generation.rs:535-558
//   build_xray_quadtree + helpers                xray/src/generation.rs:560-759
//   quadtree NodeId / Node::get_child            quadtree/src/lib.rs:57-141,143-230
//   Aabb::transform                              src/geometry/aabb.rs:58-66
//
// THIRD-PARTY, UN-VENDORED: `image::imageops::resize(.., FilterType::Lanczos3)` of the `image` crate, pinned to 0.23.10 by
// the reference's Cargo.lock.  Its source is not in /root/reference and cannot be fetched here; `resize_lanczos3` below
// restates the published algorithm of the 0.23 series (src/imageops/sample.rs: `resize` = `vertical_sample` into an
// image of the SAME pixel type, then `horizontal_sample`; f32 weights `sinc(x) sinc(x/3)` over a support of 3 x the
// down-scaling ratio; per channel `sum(v_i w_i) / sum(w_i)`, clamped to [0, 255], converted through `FloatNearest`, i.e.
// `f32::round` - the 0.23 series rounds to nearest; the truncating conversion of 0.22 and before darkened constant images).
// PARITY UNPINNED for this function: there is no golden vector of it in the reference and the crate cannot be run here.
// What IS checked: tests/test_xray_pyramid.py compares it with an independent float64 numpy evaluation of the same
// formula (equal up to the rounding boundary) and with Pillow's LANCZOS resampling (a different fixed-point
// implementation of the same filter; agreement within a few grey levels on smooth images).
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <set>
#include <vector>

#include "oracle_query.hpp"

namespace orc {

// xray_from_points with PointColorColoringStrategy (mode 1) / IntensityColoringStrategy (mode 2) and
// Binning = Some(("intensity", bin_size)): per column a map bin -> (sum, count), bin = (intensity as f64 / size) as i64
// (generation.rs:139-150); the pixel is the mean over the column's bins of the bins' means (:276-290, :339-346).
// The order in which a column's bins are summed is the iteration order of an FnvHashMap filled in batch arrival order
// (unspecified); restated as ascending bin order, results agree up to f32 rounding.
inline bool xray_tile_attr_binned(const Octree& oct, const Aabb& bbox, uint32_t w, uint32_t h, bool has_q, const Iso3& query_from_global, int mode,
                                  float p0, float p1, double bin_size, std::vector<uint8_t>& rgba) {
    const Location loc = xray_location(bbox, has_q, query_from_global);
    const size_t npix = (size_t)w * h;
    struct Col {
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        uint64_t count = 0;
    };
    std::vector<std::map<int64_t, Col>> cols(npix);
    bool seen_any = false;
    std::vector<Interval> nofilter;
    Vec3 mn = bbox.mins, dg = bbox.diag();
    for (NodeId id : nodes_in_location(oct, loc)) {
        QueryOut q;
        query_node(oct, id, loc, nofilter, q);
        const size_t n = q.src.size();
        for (size_t i = 0; i < n; ++i) {
            seen_any = true;
            Vec3 p{q.xyz[3 * i], q.xyz[3 * i + 1], q.xyz[3 * i + 2]};
            if (has_q) p = iso_transform_point(query_from_global, p);
            const uint32_t x = rust_f64_as_u32(((p.x - mn.x) / dg.x) * (double)w);
            const uint32_t y = rust_f64_as_u32((1. - ((p.y - mn.y) / dg.y)) * (double)h);
            if (!(x < w && y < h)) continue;
            const float inten = q.intensity.empty() ? 0.f : q.intensity[i];
            const int64_t bin = rust_f64_as_i64((double)inten / bin_size);
            if (mode == 2 && inten < 0.f) continue;  // see oracle_query.hpp: "negative intensities are skipped"
            Col& c = cols[(size_t)y * w + x][bin];
            if (mode == 1) {
                c.sum[0] += (float)q.rgb[3 * i] / 255.f;
                c.sum[1] += (float)q.rgb[3 * i + 1] / 255.f;
                c.sum[2] += (float)q.rgb[3 * i + 2] / 255.f;
                c.sum[3] += 255.f / 255.f;
            } else {
                c.sum[0] += inten;
            }
            c.count++;
        }
    }
    fill_transparent(rgba, npix);
    if (!seen_any) return false;
    for (size_t px = 0; px < npix; ++px) {
        if (cols[px].empty()) continue;
        uint8_t* o = &rgba[px * 4];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};  // Sum: fold(Color::default(), +) / f32 sum from 0.0
        for (const auto& kv : cols[px])
            for (int k = 0; k < 4; ++k) acc[k] += kv.second.sum[k] / (float)kv.second.count;
        const float nb = (float)cols[px].size();
        if (mode == 1) {
            for (int k = 0; k < 4; ++k) o[k] = f32_to_u8(acc[k] / nb);
        } else {
            float m = acc[0] / nb;
            m = std::fmin(std::fmax(m, p0), p1);
            const float brighten = std::log(m - p0) / std::log(p1 - p0);
            o[0] = o[1] = o[2] = f32_to_u8(brighten);
            o[3] = f32_to_u8(1.f);
        }
    }
    return true;
}

// ---- images ---------------------------------------------------------------------------------------
struct Image {  // RgbaImage: row-major, (0, 0) top left
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> px;
    bool empty() const { return px.empty(); }
};

// generation.rs:410-451.  children[i] may be empty (None); all present children are square and of one size.
inline Image build_parent(const Image* const children[4], const uint8_t bg[4]) {
    uint32_t cs = 0;
    for (int i = 0; i < 4; ++i)
        if (children[i] && !children[i]->empty()) cs = children[i]->w;
    Image out;
    out.w = out.h = cs * 2;
    out.px.resize((size_t)out.w * out.h * 4);
    for (size_t i = 0; i < (size_t)out.w * out.h; ++i) std::memcpy(&out.px[i * 4], bg, 4);
    const struct {
        int id;
        uint32_t xo, yo;
    } place[4] = {{1, 0, 0}, {0, 0, cs}, {3, cs, 0}, {2, cs, cs}};
    for (const auto& pl : place) {
        const Image* c = children[pl.id];
        if (!c || c->empty()) continue;
        for (uint32_t y = 0; y < cs; ++y) std::memcpy(&out.px[((size_t)(y + pl.yo) * out.w + pl.xo) * 4], &c->px[(size_t)y * cs * 4], (size_t)cs * 4);
    }
    return out;
}

// image 0.23 `imageops::sample`: sinc / lanczos3_kernel, f32 throughout (f32::sin = libm sinf on linux-gnu).
inline float img_sinc(float t) {
    const float a = t * 3.14159274101257324f;  // f32::consts::PI
    return t == 0.0f ? 1.0f : std::sin(a) / a;
}
inline float img_lanczos3(float x) { return std::fabs(x) < 3.0f ? img_sinc(x) * img_sinc(x / 3.0f) : 0.0f; }

// The window [left, right) and the weights of one output sample (shared by the two passes; sample.rs horizontal_sample /
// vertical_sample).  `sum` is accumulated in window order.
struct ResampleTaps {
    uint32_t left = 0;
    std::vector<float> w;
    float sum = 0.f;
};
inline ResampleTaps lanczos3_taps(uint32_t out_i, uint32_t in_size, uint32_t out_size) {
    const float ratio = (float)in_size / (float)out_size;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float src_support = 3.0f * sratio;
    float input = ((float)out_i + 0.5f) * ratio;
    int64_t left = (int64_t)std::floor(input - src_support);
    left = std::min<int64_t>(std::max<int64_t>(left, 0), (int64_t)in_size - 1);
    int64_t right = (int64_t)std::ceil(input + src_support);
    right = std::min<int64_t>(std::max<int64_t>(right, left + 1), (int64_t)in_size);
    input = input - 0.5f;
    ResampleTaps t;
    t.left = (uint32_t)left;
    for (int64_t i = left; i < right; ++i) {
        const float w = img_lanczos3(((float)i - input) / sratio);
        t.w.push_back(w);
        t.sum += w;
    }
    return t;
}
inline uint8_t img_f32_to_u8(float v) {  // NumCast::from(FloatNearest(clamp(t, 0.0, 255.0))): f32::round (half away from zero), then the cast
    const float c = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);  // image::math::utils::clamp: a NaN passes through; NumCast of NaN fails -> panic
    return (uint8_t)std::round(c);
}
inline Image resize_lanczos3(const Image& src, uint32_t nw, uint32_t nh) {
    // vertical_sample: width unchanged, height -> nh
    Image tmp;
    tmp.w = src.w;
    tmp.h = nh;
    tmp.px.resize((size_t)tmp.w * tmp.h * 4);
    for (uint32_t oy = 0; oy < nh; ++oy) {
        const ResampleTaps t = lanczos3_taps(oy, src.h, nh);
        for (uint32_t x = 0; x < src.w; ++x) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (size_t i = 0; i < t.w.size(); ++i) {
                const uint8_t* p = &src.px[((size_t)(t.left + i) * src.w + x) * 4];
                for (int k = 0; k < 4; ++k) acc[k] += (float)p[k] * t.w[i];
            }
            for (int k = 0; k < 4; ++k) tmp.px[((size_t)oy * tmp.w + x) * 4 + k] = img_f32_to_u8(acc[k] / t.sum);
        }
    }
    // horizontal_sample: height unchanged, width -> nw
    Image out;
    out.w = nw;
    out.h = nh;
    out.px.resize((size_t)nw * nh * 4);
    for (uint32_t ox = 0; ox < nw; ++ox) {
        const ResampleTaps t = lanczos3_taps(ox, tmp.w, nw);
        for (uint32_t y = 0; y < nh; ++y) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (size_t i = 0; i < t.w.size(); ++i) {
                const uint8_t* p = &tmp.px[((size_t)y * tmp.w + t.left + i) * 4];
                for (int k = 0; k < 4; ++k) acc[k] += (float)p[k] * t.w[i];
            }
            for (int k = 0; k < 4; ++k) out.px[((size_t)y * nw + ox) * 4 + k] = img_f32_to_u8(acc[k] / t.sum);
        }
    }
    return out;
}

// assign_background (generation.rs:695-720): alpha < 128 -> the background colour, every other pixel unchanged.
inline void assign_background(Image& img, const uint8_t bg[4]) {
    for (size_t i = 0; i < (size_t)img.w * img.h; ++i)
        if (img.px[i * 4 + 3] < 128) std::memcpy(&img.px[i * 4], bg, 4);
}

// ---- quadtree -------------------------------------------------------------------------------------
struct QuadId {
    uint8_t level = 0;
    uint64_t index = 0;
    bool operator<(const QuadId& o) const { return level != o.level ? level < o.level : index < o.index; }
    bool operator==(const QuadId& o) const { return level == o.level && index == o.index; }
    QuadId child(int k) const { return QuadId{(uint8_t)(level + 1), (index << 2) + (uint64_t)k}; }  // quadtree lib.rs:163-168
    QuadId parent() const { return QuadId{(uint8_t)(level - 1), index >> 2}; }                     // :178-186
};
struct QuadRect {
    double min_x = 0, min_y = 0, edge = 0;
};
inline QuadRect quad_child_rect(const QuadRect& r, int k) {  // Node::get_child, quadtree lib.rs:84-101
    const double half = r.edge / 2.;
    QuadRect c{r.min_x, r.min_y, half};
    if (k & 1) c.min_y += half;
    if (k & 2) c.min_x += half;
    return c;
}
inline QuadRect quad_rect_of(const QuadId& id, const QuadRect& root) {  // Node::from_node_id_and_root_bounding_rect, :62-82
    QuadRect r = root;
    for (int l = (int)id.level - 1; l >= 0; --l) r = quad_child_rect(r, (int)((id.index >> (2 * l)) & 3));
    return r;
}

// generation.rs:515-533
inline void find_quadtree_bounding_rect_and_levels(const Aabb& bbox, uint32_t tile_size_px, double pixel_size_m, QuadRect& rect, uint8_t& levels) {
    const double tile_size_m = (double)tile_size_px * pixel_size_m;
    levels = 0;
    double cur = tile_size_m;
    const Vec3 d = bbox.diag();
    while (cur < d.x || cur < d.y) {
        cur *= 2.;
        levels += 1;
    }
    rect = QuadRect{bbox.mins.x, bbox.mins.y, cur};
}

// Aabb::transform (aabb.rs:58-66): box of the eight transformed corners.
inline Aabb aabb_transform(const Aabb& b, const Iso3& t) {
    const Vec3 c[8] = {{b.mins.x, b.mins.y, b.mins.z}, {b.maxs.x, b.mins.y, b.mins.z}, {b.mins.x, b.maxs.y, b.mins.z}, {b.maxs.x, b.maxs.y, b.mins.z},
                       {b.mins.x, b.mins.y, b.maxs.z}, {b.maxs.x, b.mins.y, b.maxs.z}, {b.mins.x, b.maxs.y, b.maxs.z}, {b.maxs.x, b.maxs.y, b.maxs.z}};
    Vec3 p0 = iso_transform_point(t, c[0]);
    Vec3 lo = p0, hi = p0;
    for (int i = 1; i < 8; ++i) {
        const Vec3 p = iso_transform_point(t, c[i]);
        lo = {std::fmin(lo.x, p.x), std::fmin(lo.y, p.y), std::fmin(lo.z, p.z)};
        hi = {std::fmax(hi.x, p.x), std::fmax(hi.y, p.y), std::fmax(hi.z, p.z)};
    }
    return Aabb::make(lo, hi);
}

struct XrayQuadtreeParams {
    int strategy = 0;  // 0 XRay, 1 Colored, 2 ColoredWithIntensity, 3 ColoredWithHeightStddev
    float p0 = 0.f, p1 = 0.f;
    int colormap = 0;
    double bin_size = 0.0;  // > 0: Binning = Some(("intensity", bin_size)) for strategies 1 and 2
    bool has_q = false;
    Iso3 query_from_global{};
    uint8_t background[4] = {255, 255, 255, 255};
    uint32_t tile_size_px = 256;
    double pixel_size_m = 0.05;
    QuadId root{};
};
struct XrayQuadtree {
    QuadRect bounding_rect;  // of the sub-root node (Meta::bounding_rect = root_node.bounding_rect, generation.rs:609-614)
    uint8_t deepest_level = 0;
    std::map<QuadId, Image> tiles;  // Meta::nodes + the image of every node
};

// build_xray_quadtree (generation.rs:560-622) without the files: the PNGs are lossless, so the images that travel
// from level to level are what the reference re-reads from disk.
inline bool build_xray_quadtree(const Octree& oct, const XrayQuadtreeParams& pr, XrayQuadtree& out) {
    const Aabb bounding_box = pr.has_q ? aabb_transform(oct.bbox, pr.query_from_global) : oct.bbox;
    QuadRect rect;
    uint8_t deepest = 0;
    find_quadtree_bounding_rect_and_levels(bounding_box, pr.tile_size_px, pr.pixel_size_m, rect, deepest);
    if (pr.root.level > deepest) return false;  // assert "Specified root node id is outside quadtree."
    out.deepest_level = deepest;
    out.bounding_rect = quad_rect_of(pr.root, rect);
    // get_nodes_at_level (:535-551)
    std::vector<QuadId> leaves{pr.root};
    for (int l = pr.root.level; l < deepest; ++l) {
        std::vector<QuadId> next;
        for (const QuadId& n : leaves)
            for (int k = 0; k < 4; ++k) next.push_back(n.child(k));
        leaves.swap(next);
    }
    // create_leaf_nodes (:624-667)
    std::set<QuadId> current;
    for (const QuadId& id : leaves) {
        const QuadRect r = quad_rect_of(id, rect);
        const Aabb bb = Aabb::make({r.min_x, r.min_y, bounding_box.mins.z}, {r.min_x + r.edge, r.min_y + r.edge, bounding_box.maxs.z});
        Image img;
        img.w = img.h = pr.tile_size_px;
        bool any;
        if (pr.strategy == 0)
            any = xray_tile(oct, bb, img.w, img.h, pr.has_q, pr.query_from_global, img.px, nullptr);
        else if (pr.bin_size > 0.0 && (pr.strategy == 1 || pr.strategy == 2))
            any = xray_tile_attr_binned(oct, bb, img.w, img.h, pr.has_q, pr.query_from_global, pr.strategy, pr.p0, pr.p1, pr.bin_size, img.px);
        else
            any = xray_tile_attr(oct, bb, img.w, img.h, pr.has_q, pr.query_from_global, pr.strategy, pr.p0, pr.p1, pr.colormap, img.px);
        if (!any) continue;
        assign_background(img, pr.background);  // :695-720, applied to the created leaves only
        out.tiles[id] = std::move(img);
        current.insert(id);
    }
    // create_non_leaf_nodes (:669-693) + build_node (:722-759)
    for (int level = (int)deepest - 1; level >= (int)pr.root.level; --level) {
        std::set<QuadId> parents;
        for (const QuadId& id : current) parents.insert(id.parent());
        for (const QuadId& id : parents) {
            const Image* ch[4];
            for (int k = 0; k < 4; ++k) {
                auto it = out.tiles.find(id.child(k));
                ch[k] = it == out.tiles.end() ? nullptr : &it->second;
            }
            const Image large = build_parent(ch, pr.background);
            out.tiles[id] = resize_lanczos3(large, pr.tile_size_px, pr.tile_size_px);
        }
        current.swap(parents);
    }
    return true;
}

}  // namespace orc
