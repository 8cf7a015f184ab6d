// xray_api.inl — the C ABI of the X-ray pipeline beyond the leaf tile (included by pcv_api.cu after
// query_api.inl): binned columns, assign_background, build_parent + Lanczos3 reduction, and build_xray_quadtree as one call
// that keeps every tile in HBM from the leaves to the root and hands each finished image to the caller (PNG encoding and
// meta.pb stay on the host, as SURVEY 8 f3 says).
//
// Reference: xray/src/generation.rs:129-157 (bins), :410-451 (build_parent), :515-558 (rect, levels, leaves, bounding box),
// :560-622 (build_xray_quadtree), :624-667 (create_leaf_nodes), :669-693 (create_non_leaf_nodes), :695-720
// (assign_background), :722-759 (build_node).

#include "xray_png.hpp"

namespace {

using namespace pcv;

inline uint32_t pack_rgba(const uint8_t c[4]) { return (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24); }

struct DevTaps {
    ResampleTaps t{};
};
DevTaps upload_taps(Scratch& s, const ResampleTable& tb) {
    DevTaps d;
    d.t.left = s.upload(tb.left.data(), tb.left.size());
    d.t.first = s.upload(tb.first.data(), tb.first.size());
    d.t.count = s.upload(tb.count.data(), tb.count.size());
    d.t.sum = s.upload(tb.sum.data(), tb.sum.size());
    d.t.w = s.upload(tb.w.data(), tb.w.size());
    return d;
}

// One parent tile on the device: children (device pointers, cs x cs RGBA, or null) -> d_out (tile_px x tile_px RGBA).
// d_tmp holds the vertically reduced mosaic: (2 cs) x tile_px pixels.
void parent_tile_device(pcv_ctx* c, const uint8_t* const d_children[4], uint32_t cs, uint32_t bg, uint32_t tile_px, const DevTaps& tv, const DevTaps& th,
                        uint32_t* d_tmp, uint32_t* d_out) {
    ResampleVArgs v{};
    for (int k = 0; k < 4; ++k) v.src.child[k] = d_children[k];
    v.src.cs = cs;
    v.src.bg = bg;
    v.taps = tv.t;
    v.in_w = 2 * cs;
    v.out_h = tile_px;
    v.out = d_tmp;
    const size_t nv = (size_t)v.in_w * v.out_h, nh = (size_t)tile_px * tile_px;
    const uint32_t cap = (uint32_t)c->sm_count * 16;
    k_xray_resample_v<<<(uint32_t)std::min<size_t>((nv + 255) / 256, cap), 256, 0, c->stream>>>(v);
    ResampleHArgs hh{};
    hh.in = d_tmp;
    hh.taps = th.t;
    hh.in_w = 2 * cs;
    hh.out_w = tile_px;
    hh.h = tile_px;
    hh.out = d_out;
    k_xray_resample_h<<<(uint32_t)std::min<size_t>((nh + 255) / 256, cap), 256, 0, c->stream>>>(hh);
    c->be->launches += 2;
    CU(cudaGetLastError());
}

}  // namespace

extern "C" {

int pcv_xray_tile_attr_binned(const pcv_octree* oc, const double tmin[3], const double tmax[3], uint32_t w, uint32_t h, const double* qfg, int mode,
                              float p0, float p1, double bin_size, uint8_t* rgba_out, int* any_out) {
    if (!oc || !tmin || !tmax || !rgba_out || w == 0 || h == 0) return fail(PCV_ERR_INVALID, "null argument or empty image");
    if (bin_size == 0.0) return fail(PCV_ERR_INVALID, "bin size 0 (use pcv_xray_tile_attr for Binning = None)");
    if ((uint64_t)w * h > 0xFFFFFFFFull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-1 pixels in one binned X-ray tile");
    if (int rc = xray_attr_check(oc, mode, bin_size)) return rc;
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    return xray_tile_attr_core(o, tmin, tmax, w, h, qfg, mode, p0, p1, 0, bin_size, rgba_out, any_out, nullptr);
    API_CATCH
}

int pcv_xray_assign_background(pcv_ctx* c, uint8_t* rgba, uint64_t npix, const uint8_t background[4]) {
    if (!c || (!rgba && npix) || !background) return fail(PCV_ERR_INVALID, "null argument");
    if (npix == 0) return PCV_OK;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    Scratch s(c);
    uint32_t* d = (uint32_t*)s.alloc<uint8_t>(npix * 4);
    c->be->h2d(d, rgba, npix * 4);
    k_xray_background<<<(uint32_t)std::min<uint64_t>((npix + 255) / 256, (uint64_t)c->sm_count * 16), 256, 0, c->stream>>>(d, npix, pack_rgba(background));
    c->be->launches++;
    CU(cudaGetLastError());
    c->be->d2h(rgba, d, npix * 4);
    return PCV_OK;
    API_CATCH
}

int pcv_xray_build_parent(pcv_ctx* c, const uint8_t* const children[4], uint32_t child_px, const uint8_t background[4], uint32_t tile_px, uint8_t* rgba_out) {
    if (!c || !children || !background || !rgba_out || child_px == 0 || tile_px == 0) return fail(PCV_ERR_INVALID, "null argument or empty image");
    if (!children[0] && !children[1] && !children[2] && !children[3]) return fail(PCV_ERR_INVALID, "No children passed to 'build_parent'.");
    if (child_px > 32768 || tile_px > 65535) return fail(PCV_ERR_UNSUPPORTED, "tile too large");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    Scratch s(c);
    const size_t cbytes = (size_t)child_px * child_px * 4;
    const uint8_t* dch[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k)
        if (children[k]) {
            uint8_t* d = s.alloc<uint8_t>(cbytes);
            c->be->h2d(d, children[k], cbytes);
            dch[k] = d;
        }
    const DevTaps tv = upload_taps(s, make_lanczos3_table(2 * child_px, tile_px));
    const DevTaps th = tv;  // the mosaic is square: both axes reduce 2 child_px -> tile_px with the same taps
    uint32_t* dtmp = s.alloc<uint32_t>((size_t)2 * child_px * tile_px);
    uint32_t* dout = s.alloc<uint32_t>((size_t)tile_px * tile_px);
    parent_tile_device(c, dch, child_px, pack_rgba(background), tile_px, tv, th, dtmp, dout);
    c->be->d2h(rgba_out, dout, (size_t)tile_px * tile_px * 4);
    return PCV_OK;
    API_CATCH
}

int pcv_xray_quadtree(const pcv_octree* oc, const pcv_xray_quadtree_params* pr, pcv_xray_tile_fn on_tile, void* user, pcv_xray_quadtree_info* info) {
    if (!oc || !pr || !info) return fail(PCV_ERR_INVALID, "null argument");
    if (pr->strategy < 0 || pr->strategy > PCV_XRAY_HEIGHT_STDDEV) return fail(PCV_ERR_INVALID, "unknown colouring strategy %d", pr->strategy);
    if (pr->strategy != 0)
        if (int rc = xray_attr_check(oc, pr->strategy, pr->bin_size)) return rc;
    if (pr->tile_size_px == 0 || pr->tile_size_px > 32768) return fail(PCV_ERR_INVALID, "tile size %u", pr->tile_size_px);
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    *info = pcv_xray_quadtree_info{};
    const double* qfg = pr->has_query_from_global ? pr->query_from_global : nullptr;
    // get_bounding_box (:553-558): the octree's box, or Aabb::transform of it (aabb.rs:58-66: box of the transformed corners)
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; ++a) bmin[a] = o->bbox_min[a], bmax[a] = o->bbox_max[a];
    if (qfg) {
        double lo[3], hi[3];
        for (int k = 0; k < 8; ++k) {  // compute_corners order (aabb.rs:125-138): x fastest, then y, then z
            const V3 p = iso_apply(qfg, V3{(k & 1) ? bmax[0] : bmin[0], (k & 2) ? bmax[1] : bmin[1], (k & 4) ? bmax[2] : bmin[2]});
            const double v[3] = {p.x, p.y, p.z};
            for (int a = 0; a < 3; ++a) {
                lo[a] = k == 0 ? v[a] : std::fmin(lo[a], v[a]);
                hi[a] = k == 0 ? v[a] : std::fmax(hi[a], v[a]);
            }
        }
        for (int a = 0; a < 3; ++a) bmin[a] = lo[a], bmax[a] = hi[a];
    }
    QuadRect rect{};
    uint8_t deepest = 0;
    if (!quadtree_rect_and_levels(bmin, bmax, pr->tile_size_px, pr->pixel_size_m, rect, deepest))
        return fail(PCV_ERR_INVALID, "pixel size %g does not give a finite quadtree", pr->pixel_size_m);
    if (pr->root_level > deepest) return fail(PCV_ERR_INVALID, "Specified root node id is outside quadtree.");
    if (pr->root_level < 32 && (pr->root_index >> (2 * pr->root_level)) != 0) return fail(PCV_ERR_INVALID, "root node index outside its level");
    if (deepest - pr->root_level > 12) return fail(PCV_ERR_UNSUPPORTED, "more than 4^12 leaf tiles below the root node");
    const QuadId root{pr->root_level, pr->root_index};
    const QuadRect root_rect = quad_rect_of(root, rect);
    info->rect_min_x = root_rect.min_x;
    info->rect_min_y = root_rect.min_y;
    info->rect_edge = root_rect.edge;
    info->deepest_level = deepest;
    info->tile_size_px = pr->tile_size_px;
    const uint32_t T = pr->tile_size_px;
    const size_t tbytes = (size_t)T * T * 4;
    const uint32_t bg = pack_rgba(pr->background);
    std::vector<uint8_t> host(on_tile ? tbytes : 0);
    std::map<QuadId, uint8_t*> level_tiles;  // the level being consumed (device images)
    struct Free {
        pcv_ctx* c;
        std::map<QuadId, uint8_t*>* a;
        std::map<QuadId, uint8_t*>* b;
        ~Free() {
            for (auto* m : {a, b})
                if (m)
                    for (auto& kv : *m) c->be->dfree(kv.second);
        }
    };
    std::map<QuadId, uint8_t*> next_tiles;
    Free guard{c, &level_tiles, &next_tiles};
    struct Ev {
        cudaEvent_t e = nullptr;
        ~Ev() {
            if (e) cudaEventDestroy(e);
        }
    } ev0, ev1, ev2;
    CU(cudaEventCreate(&ev0.e));
    CU(cudaEventCreate(&ev1.e));
    CU(cudaEventCreate(&ev2.e));
    const cudaEvent_t e0 = ev0.e, e1 = ev1.e, e2 = ev2.e;
    const uint64_t l0 = c->be->launches;
    auto deliver = [&](const QuadId& id, const uint8_t* d) -> int {
        info->num_nodes++;
        if (!on_tile) return 0;
        c->be->d2h(host.data(), d, tbytes);
        return on_tile(user, id.level, id.index, host.data(), T);
    };
    // ---- leaves: get_nodes_at_level + create_leaf_nodes + assign_background --------------------------------------
    CU(cudaEventRecord(e0, c->stream));
    const uint64_t nleaf = 1ull << (2 * (deepest - pr->root_level));
    uint8_t* spare = nullptr;
    uint64_t leaf_points = 0;
    for (uint64_t k = 0; k < nleaf; ++k) {
        const QuadId id{deepest, (pr->root_index << (2 * (deepest - pr->root_level))) + k};
        const QuadRect r = quad_rect_of(id, rect);
        const double tmin[3] = {r.min_x, r.min_y, bmin[2]};
        const double r_max_x = r.min_x + r.edge, r_max_y = r.min_y + r.edge;  // Rect::max (quadtree lib.rs:39-41)
        const double tmax[3] = {r_max_x, r_max_y, bmax[2]};
        if (!spare) spare = (uint8_t*)c->be->dmalloc(tbytes);
        int any = 0, rc;
        if (pr->strategy == 0) {
            rc = xray_tile_core(o, tmin, tmax, T, T, qfg, nullptr, nullptr, &any, spare);
            leaf_points += c->xstats.points;
        } else {
            rc = xray_tile_attr_core(o, tmin, tmax, T, T, qfg, pr->strategy, pr->p0, pr->p1, pr->colormap, pr->bin_size, nullptr, &any, spare);
        }
        if (rc != PCV_OK) {
            c->be->dfree(spare);
            return rc;
        }
        if (!any) continue;  // xray_from_points returned None: no image, the node does not exist
        k_xray_background<<<(uint32_t)std::min<size_t>(((size_t)T * T + 255) / 256, (size_t)c->sm_count * 16), 256, 0, c->stream>>>((uint32_t*)spare, (size_t)T * T, bg);
        c->be->launches++;
        level_tiles[id] = spare;
        spare = nullptr;
        info->num_leaves++;
    }
    if (spare) c->be->dfree(spare);
    CU(cudaGetLastError());
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&info->ms_leaves, e0, e1);
    for (auto& kv : level_tiles)
        if (int stop = deliver(kv.first, kv.second)) return fail(PCV_ERR_CANCELLED, "X-ray quadtree: the tile callback returned %d", stop);
    // ---- parents: create_non_leaf_nodes + build_node -----------------------------------------------------------------
    Scratch s(c);
    DevTaps taps;
    uint32_t* dtmp = nullptr;
    if (deepest > pr->root_level && !level_tiles.empty()) {
        taps = upload_taps(s, make_lanczos3_table(2 * T, T));
        dtmp = s.alloc<uint32_t>((size_t)2 * T * T);
    }
    for (int level = (int)deepest - 1; level >= (int)pr->root_level && !level_tiles.empty(); --level) {
        CU(cudaEventRecord(e1, c->stream));
        for (auto it = level_tiles.begin(); it != level_tiles.end();) {
            const QuadId pid = quad_parent(it->first);
            const uint8_t* ch[4] = {nullptr, nullptr, nullptr, nullptr};
            // the map is ordered by (level, index): the children of one parent are adjacent
            while (it != level_tiles.end() && quad_parent(it->first).index == pid.index) {
                ch[it->first.index & 3] = it->second;
                ++it;
            }
            uint8_t* dout = (uint8_t*)c->be->dmalloc(tbytes);
            next_tiles[pid] = dout;
            parent_tile_device(c, ch, T, bg, T, taps, taps, dtmp, (uint32_t*)dout);
        }
        CU(cudaEventRecord(e2, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e1, e2);
        info->ms_parents += ms;
        for (auto& kv : next_tiles)
            if (int stop = deliver(kv.first, kv.second)) return fail(PCV_ERR_CANCELLED, "X-ray quadtree: the tile callback returned %d", stop);
        for (auto& kv : level_tiles) c->be->dfree(kv.second);
        level_tiles.clear();
        level_tiles.swap(next_tiles);
    }
    info->kernel_launches = (uint32_t)(c->be->launches - l0);
    info->leaf_points = leaf_points;
    return PCV_OK;
    API_CATCH
}

// build_xray_quadtree with the reference's outputs: <directory>/<node id>.png for every tile and the quadtree's meta file
// (generation.rs:560-622: create_dir, images, meta.to_disk(get_meta_pb_path(output_directory, root_node_id))).
int pcv_xray_quadtree_write_dir(const pcv_octree* oc, const pcv_xray_quadtree_params* pr, const char* dir, pcv_xray_quadtree_info* info) {
    if (!oc || !pr || !dir || !info) return fail(PCV_ERR_INVALID, "null argument");
    mkdir(dir, 0777);  // "Ignore errors, maybe directory is already there." (:565-566)
    struct State {
        std::string base;
        XrayMetaData meta;
        std::string err;
    } st;
    st.base = std::string(dir) + "/";
    auto on_tile = [](void* user, uint8_t level, uint64_t index, const uint8_t* rgba, uint32_t t) -> int {
        State* s = (State*)user;
        std::string png;
        const std::string path = s->base + quad_node_name(level, index) + ".png";
        if (!encode_png_rgba(rgba, t, t, png) || !write_whole_file(path, png.data(), png.size())) {
            s->err = "cannot write " + path;
            return 1;
        }
        s->meta.nodes.emplace_back((uint32_t)level, index);
        return 0;
    };
    const int rc = pcv_xray_quadtree(oc, pr, on_tile, &st, info);
    if (rc == PCV_ERR_CANCELLED && !st.err.empty()) return fail(PCV_ERR_IO, "%s", st.err.c_str());
    if (rc != PCV_OK) return rc;
    st.meta.min_x = info->rect_min_x;
    st.meta.min_y = info->rect_min_y;
    st.meta.edge = info->rect_edge;
    st.meta.deepest_level = info->deepest_level;
    st.meta.tile_size = info->tile_size_px;
    const std::string name = quad_node_name(pr->root_level, pr->root_index);  // "r..." -> "meta..." (utils.rs:7-11)
    const std::string path = st.base + "meta" + name.substr(1) + ".pb";
    const std::string buf = encode_xray_meta(st.meta);
    if (!write_whole_file(path, buf.data(), buf.size())) return fail(PCV_ERR_IO, "cannot write %s", path.c_str());
    return PCV_OK;
}

}  // extern "C"
