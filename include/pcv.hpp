// pcv.hpp — header-only C++ host layer over the C ABI (pcv.h), mirroring the names, argument meaning and error behaviour
// of the reference's Rust interface for this path, so that callers (and tests) read like the reference's own:
//
//   point_viewer::octree::build_octree                  src/octree/generation.rs:289-295   -> pcv::build_octree
//   point_viewer::octree::Octree::{from_data_provider, get_visible_nodes, get_node_data, nodes_in_location}
//                                                       src/octree/mod.rs:156,228,285,329  -> pcv::Octree
//   point_viewer::iterator::{PointQuery, PointLocation, ParallelIterator::try_for_each_batch}
//                                                       src/iterator.rs:13-20,66-72,255    -> pcv::PointQuery, pcv::ParallelIterator
//   point_viewer::{PointsBatch, NodeId}                 src/lib.rs:102-107, src/octree/node.rs:52-111
//   point_viewer::s2_cells::S2Cells, read_write::S2Splitter   src/s2_cells/mod.rs, src/read_write/s2.rs  -> pcv::S2Cells, pcv::s2_split
//
// Error behaviour: the reference panics (unwrap) in build_octree / get_visible_nodes and returns Result elsewhere; here
// every failure is a pcv::Error exception carrying the pcv_status and text (callers that want the panic semantics let it
// propagate).  A consumer callback returning false cancels the stream (== Err -> ErrorKind::Channel).
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "pcv.h"

namespace pcv {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};
inline void check(int rc) {
    if (rc != PCV_OK) throw Error(rc, pcv_last_error());
}

// NodeId: u128 = level << 120 | index (node.rs:52-111), Display "r" + octal path (node.rs:73-86)
struct NodeId {
    uint64_t high = 0, low = 0;
    int level() const { return (int)(high >> 56); }
    std::string to_string() const {
        unsigned __int128 v = ((unsigned __int128)high << 64) | low;
        std::string s(1, 'r');
        for (int i = level() - 1; i >= 0; --i) s.push_back((char)('0' + (int)((v >> (3 * i)) & 7)));
        return s;
    }
    bool operator==(const NodeId& o) const { return high == o.high && low == o.low; }
};

struct Aabb {  // Aabb::new takes inf/sup of the two corners (aabb.rs:19-24)
    std::array<double, 3> min, max;
    Aabb(std::array<double, 3> a, std::array<double, 3> b) {
        for (int i = 0; i < 3; ++i) {
            min[i] = a[i] < b[i] ? a[i] : b[i];
            max[i] = a[i] < b[i] ? b[i] : a[i];
        }
    }
};

// PointsBatch (lib.rs:102-107): AoS positions + colour (U8Vec3) and optional intensity (F32)
struct PointsBatch {
    std::vector<std::array<double, 3>> position;
    std::vector<std::array<uint8_t, 3>> color;
    std::vector<float> intensity;         // empty if absent
    std::vector<uint64_t> source_index;   // provenance (not in the reference)
};

// PointLocation (iterator.rs:13-20).  Frustum / Obb carry the fields the reference structs hold.
struct PointLocation {
    pcv_location raw{};
    static PointLocation AllPoints() {
        PointLocation l;
        l.raw.kind = PCV_LOC_ALL;
        return l;
    }
    static PointLocation from(const Aabb& b) {
        PointLocation l;
        l.raw.kind = PCV_LOC_AABB;
        for (int i = 0; i < 3; ++i) {
            l.raw.aabb_min[i] = b.min[i];
            l.raw.aabb_max[i] = b.max[i];
        }
        return l;
    }
    // Frustum{query_from_clip, clip_from_query}: column-major 4x4 (nalgebra storage)
    static PointLocation Frustum(const double clip_from_query[16], const double query_from_clip[16]) {
        PointLocation l;
        l.raw.kind = PCV_LOC_FRUSTUM;
        for (int i = 0; i < 16; ++i) {
            l.raw.clip_from_query[i] = clip_from_query[i];
            l.raw.query_from_clip[i] = query_from_clip[i];
        }
        return l;
    }
    // Obb{query_from_obb, obb_from_query, half_extent}: isometries as tx,ty,tz,qi,qj,qk,qw
    static PointLocation Obb(const double query_from_obb[7], const double obb_from_query[7], const double half_extent[3]) {
        PointLocation l;
        l.raw.kind = PCV_LOC_OBB;
        for (int i = 0; i < 7; ++i) {
            l.raw.query_from_obb[i] = query_from_obb[i];
            l.raw.obb_from_query[i] = obb_from_query[i];
        }
        for (int i = 0; i < 3; ++i) l.raw.half_extent[i] = half_extent[i];
        return l;
    }
};

struct ClosedInterval {
    double lower_bound, upper_bound;
};
struct PointQuery {  // iterator.rs:66-72 (attributes: colour is always delivered, intensity when the octree has it)
    PointLocation location = PointLocation::AllPoints();
    std::vector<ClosedInterval> filter_intervals;  // on "intensity"
};

class Context {
   public:
    explicit Context(int device = 0, uint64_t max_points_per_node = 0) {
        pcv_config cfg{max_points_per_node, 0, 0};
        check(pcv_create(device, &cfg, &h_));
    }
    ~Context() { pcv_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    pcv_ctx* raw() const { return h_; }

   private:
    pcv_ctx* h_ = nullptr;
};

struct NodeData {  // octree/mod.rs:147-152
    pcv_node_meta meta;
    std::vector<uint8_t> position, color;
};

class Octree {
   public:
    Octree(pcv_octree* o) : o_(o) { load_table(); }
    // Octree::from_data_provider(OnDiskDataProvider{directory})
    static Octree from_directory(Context& ctx, const std::string& dir) {
        pcv_octree* o = nullptr;
        check(pcv_octree_load_dir(ctx.raw(), dir.c_str(), &o));
        return Octree(o);
    }
    Octree(Octree&& other) noexcept : o_(other.o_), nodes_(std::move(other.nodes_)) { other.o_ = nullptr; }
    Octree(const Octree&) = delete;
    ~Octree() {
        if (o_) pcv_octree_free(o_);
    }
    const std::vector<pcv_node_meta>& nodes() const { return nodes_; }
    int64_t num_points() const {
        int64_t n = 0;
        for (auto& m : nodes_) n += m.num_points;
        return n;
    }
    std::vector<NodeId> get_visible_nodes(const double projection_matrix[16]) const {  // mod.rs:228 (panics if singular)
        std::vector<uint64_t> ids(2 * nodes_.size() + 2);
        uint64_t n = 0;
        check(pcv_visible_nodes(o_, projection_matrix, ids.data(), nodes_.size(), &n));
        return to_ids(ids, n);
    }
    std::vector<NodeId> nodes_in_location(const PointLocation& loc) const {  // mod.rs:329-331
        std::vector<uint64_t> ids(2 * nodes_.size() + 2);
        uint64_t n = 0;
        check(pcv_nodes_in_location(o_, &loc.raw, ids.data(), nodes_.size(), &n));
        return to_ids(ids, n);
    }
    NodeData get_node_data(const NodeId& id) const {  // mod.rs:285-307
        for (auto& m : nodes_)
            if (m.id_high == id.high && m.id_low == id.low) {
                NodeData d;
                d.meta = m;
                const size_t bpc = m.position_encoding == 1 ? 1 : m.position_encoding == 2 ? 2 : m.position_encoding == 3 ? 4 : 8;
                d.position.resize((size_t)m.num_points * 3 * bpc);
                d.color.resize((size_t)m.num_points * 3);
                check(pcv_octree_node_data(o_, id.high, id.low, d.position.data(), d.color.data(), nullptr, nullptr));
                return d;
            }
        throw Error(PCV_ERR_NOT_FOUND, "node " + id.to_string() + " not found");
    }
    // The web viewer's /nodes_data reply for a list of nodes (octree_web_viewer/src/backend.rs:92-165), gathered on the GPU.
    std::vector<uint8_t> nodes_data_blob(const std::vector<NodeId>& ids) const {
        std::vector<uint64_t> hl;
        for (auto& id : ids) hl.push_back(id.high), hl.push_back(id.low);
        uint64_t size = 0;
        check(pcv_nodes_data_blob(o_, hl.data(), (uint32_t)ids.size(), nullptr, 0, &size));
        std::vector<uint8_t> blob(size);
        if (size) check(pcv_nodes_data_blob(o_, hl.data(), (uint32_t)ids.size(), blob.data(), size, &size));
        return blob;
    }
    // xray_from_points for one tile with ColoringStrategyKind::{Colored, ColoredWithIntensity, ColoredWithHeightStddev}
    // (xray/src/generation.rs:76-97); returns false for a tile without points (None in the reference).
    bool xray_tile_attr(const Aabb& tile, uint32_t w, uint32_t h, int strategy, float p0, float p1, int colormap, std::vector<uint8_t>& rgba,
                        const double* query_from_global7 = nullptr) const {
        rgba.assign((size_t)w * h * 4, 0);
        int any = 0;
        check(pcv_xray_tile_attr(o_, tile.min.data(), tile.max.data(), w, h, query_from_global7, strategy, p0, p1, colormap, rgba.data(), &any));
        return any != 0;
    }
    // ... with Binning = Some(("intensity", bin_size)) for the Colored / ColoredWithIntensity strategies (generation.rs:129-157)
    bool xray_tile_attr_binned(const Aabb& tile, uint32_t w, uint32_t h, int strategy, float p0, float p1, double bin_size, std::vector<uint8_t>& rgba,
                               const double* query_from_global7 = nullptr) const {
        rgba.assign((size_t)w * h * 4, 0);
        int any = 0;
        check(pcv_xray_tile_attr_binned(o_, tile.min.data(), tile.max.data(), w, h, query_from_global7, strategy, p0, p1, bin_size, rgba.data(), &any));
        return any != 0;
    }
    // build_xray_quadtree (xray/src/generation.rs:560-622): every tile of the quadtree, leaves to root, through `on_tile`
    // (level, index, RGBA tile_size_px^2); returns what the reference writes into the quadtree's meta.pb.
    template <class F>
    pcv_xray_quadtree_info build_xray_quadtree(const pcv_xray_quadtree_params& params, F&& on_tile) const {
        struct Thunk {
            F* f;
            static int call(void* user, uint8_t level, uint64_t index, const uint8_t* rgba, uint32_t tile_px) {
                (*static_cast<Thunk*>(user)->f)(level, index, rgba, tile_px);
                return 0;
            }
        } th{&on_tile};
        pcv_xray_quadtree_info info{};
        check(pcv_xray_quadtree(o_, &params, &Thunk::call, &th, &info));
        return info;
    }
    void write_to_directory(const std::string& dir) const { check(pcv_octree_write_dir(o_, dir.c_str())); }
    pcv_octree* raw() const { return o_; }

   private:
    void load_table() {
        uint64_t nn = 0;
        check(pcv_octree_info(o_, &nn, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        nodes_.resize(nn);
        check(pcv_octree_nodes(o_, nodes_.data(), nn));
    }
    static std::vector<NodeId> to_ids(const std::vector<uint64_t>& v, uint64_t n) {
        std::vector<NodeId> out(n);
        for (uint64_t i = 0; i < n; ++i) out[i] = NodeId{v[2 * i], v[2 * i + 1]};
        return out;
    }
    pcv_octree* o_;
    std::vector<pcv_node_meta> nodes_;
};

// build_octree(output_directory, resolution, bounding_box, input, attributes) — generation.rs:289-295.  `input` is drained on
// the calling thread; colour is mandatory; attributes selects whether intensity is carried.  Returns the GPU-resident
// octree as well (the reference returns ()).
template <class BatchIterator>
inline Octree build_octree(Context& ctx, const std::string& output_directory, double resolution, const Aabb& bounding_box, BatchIterator begin,
                           BatchIterator end, const std::vector<std::string>& attributes = {"color"}) {
    std::vector<std::array<double, 3>> pos;
    std::vector<std::array<uint8_t, 3>> col;
    std::vector<float> inten;
    bool want_i = false;
    for (auto& a : attributes) want_i = want_i || a == "intensity";
    for (BatchIterator it = begin; it != end; ++it) {
        const PointsBatch& b = *it;
        if (b.color.size() != b.position.size()) throw Error(PCV_ERR_INVALID, "color is mandatory (on_disk.rs:23-33)");
        pos.insert(pos.end(), b.position.begin(), b.position.end());
        col.insert(col.end(), b.color.begin(), b.color.end());
        if (want_i && !b.intensity.empty()) inten.insert(inten.end(), b.intensity.begin(), b.intensity.end());
    }
    pcv_points pts{};
    const double* base = pos.empty() ? nullptr : pos[0].data();
    pts.x = base;
    pts.y = base ? base + 1 : nullptr;
    pts.z = base ? base + 2 : nullptr;
    pts.stride = 3;  // AoS Point3<f64>, no transpose
    pts.rgb = col.empty() ? nullptr : col[0].data();
    pts.intensity = (want_i && inten.size() == pos.size() && !inten.empty()) ? inten.data() : nullptr;
    pts.n = pos.size();
    pcv_octree* o = nullptr;
    check(pcv_build_octree(ctx.raw(), &pts, resolution, bounding_box.min.data(), bounding_box.max.data(), &o));
    Octree tree(o);
    if (!output_directory.empty()) tree.write_to_directory(output_directory);
    return tree;
}

// build_octree_from_file(output_directory, resolution, filename, attributes) — generation.rs:272-287: the PLY body goes to the
// GPU as raw records, the bounding-box pass is fused into the unpack kernel.
inline Octree build_octree_from_file(Context& ctx, const std::string& output_directory, double resolution, const std::string& filename,
                                     const std::vector<std::string>& attributes = {"color"}) {
    bool want_i = false;
    for (auto& a : attributes) want_i = want_i || a == "intensity";
    pcv_octree* o = nullptr;
    check(pcv_build_octree_from_file(ctx.raw(), filename.c_str(), resolution, want_i ? 1 : 0, &o));
    Octree tree(o);
    if (!output_directory.empty()) tree.write_to_directory(output_directory);
    return tree;
}

// ---- the S2-cell point cloud: point_viewer::s2_cells::S2Cells + read_write::S2Splitter (src/s2_cells/mod.rs, src/read_write/s2.rs) ----
using CellID = uint64_t;                 // s2::cellid::CellID(u64)
using CellUnion = std::vector<CellID>;   // s2::cellunion::CellUnion(Vec<CellID>)

// CellID::to_token (the per-cell file stem of the S2 directory layout)
inline std::string cell_token(CellID id) {
    if (id == 0) return "X";
    static const char* hex = "0123456789abcdef";
    std::string s;
    for (int k = 15; k >= 0; --k) s.push_back(hex[(id >> (4 * k)) & 15]);
    while (!s.empty() && s.back() == '0') s.pop_back();
    return s;
}

class S2Cells {
   public:
    explicit S2Cells(pcv_s2cloud* s) : s_(s) {}
    // S2Cells::from_data_provider over an on-disk directory (mod.rs:203-216)
    static S2Cells from_directory(Context& ctx, const std::string& directory) {
        pcv_s2cloud* s = nullptr;
        check(pcv_s2_load_dir(ctx.raw(), directory.c_str(), &s));
        return S2Cells(s);
    }
    ~S2Cells() {
        if (s_) pcv_s2_free(s_);
    }
    S2Cells(S2Cells&& o) noexcept : s_(o.s_) { o.s_ = nullptr; }
    S2Cells(const S2Cells&) = delete;
    S2Cells& operator=(const S2Cells&) = delete;

    uint64_t num_points() const {
        uint64_t n = 0;
        check(pcv_s2_info(s_, nullptr, &n, nullptr, nullptr, nullptr, nullptr, nullptr));
        return n;
    }
    Aabb bounding_box() const {  // PointCloud::bounding_box (mod.rs:192-194)
        std::array<double, 3> mn{}, mx{};
        check(pcv_s2_info(s_, nullptr, nullptr, nullptr, mn.data(), mx.data(), nullptr, nullptr));
        return Aabb(mn, mx);
    }
    // S2Meta::get_cells: (cell id, num_points), in id order
    std::vector<std::pair<CellID, uint64_t>> cells() const {
        uint64_t nc = 0;
        check(pcv_s2_info(s_, &nc, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        std::vector<uint64_t> ids(nc), cnt(nc);
        check(pcv_s2_cells(s_, ids.data(), cnt.data()));
        std::vector<std::pair<CellID, uint64_t>> out(nc);
        for (uint64_t k = 0; k < nc; ++k) out[k] = {ids[k], cnt[k]};
        return out;
    }
    // PointCloud::nodes_in_location for PointLocation::AllPoints (nullptr) and PointLocation::S2Cells (mod.rs:157-168)
    std::vector<CellID> nodes_in_location(const CellUnion* cell_union) const {
        uint64_t n = 0;
        const uint64_t* u = cell_union ? cell_union->data() : nullptr;
        const uint32_t nu = cell_union ? (uint32_t)cell_union->size() : 0;
        check(pcv_s2_cells_in_union(s_, u, nu, nullptr, 0, &n));
        std::vector<CellID> out(n);
        check(pcv_s2_cells_in_union(s_, u, nu, out.data(), n, &n));
        return out;
    }
    // points_in_node (mod.rs:174-190) as one batch
    PointsBatch points_in_node(CellID id) const {
        PointsBatch b;
        for (auto& c : cells())
            if (c.first == id) {
                int hc = 0, hi = 0;
                check(pcv_s2_info(s_, nullptr, nullptr, nullptr, nullptr, nullptr, &hc, &hi));
                b.position.resize(c.second);
                if (hc) b.color.resize(c.second);
                if (hi) b.intensity.resize(c.second);
                check(pcv_s2_cell_data(s_, id, c.second ? b.position[0].data() : nullptr, hc && c.second ? b.color[0].data() : nullptr,
                                       hi && c.second ? b.intensity.data() : nullptr, nullptr));
                return b;
            }
        throw Error(PCV_ERR_NOT_FOUND, "cell " + cell_token(id) + " not found");
    }
    // the filtered point stream of PointLocation::AllPoints / S2Cells (iterator.rs:96-119 with the CellUnion as PointCulling)
    PointsBatch query(const CellUnion* cell_union) const {
        const uint64_t* u = cell_union ? cell_union->data() : nullptr;
        const uint32_t nu = cell_union ? (uint32_t)cell_union->size() : 0;
        uint64_t n = 0;
        check(pcv_s2_query_union(s_, u, nu, nullptr, nullptr, nullptr, nullptr, 0, &n, nullptr));
        int hc = 0, hi = 0;
        check(pcv_s2_info(s_, nullptr, nullptr, nullptr, nullptr, nullptr, &hc, &hi));
        PointsBatch b;
        b.position.resize(n);
        if (hc) b.color.resize(n);
        if (hi) b.intensity.resize(n);
        if (n)
            check(pcv_s2_query_union(s_, u, nu, b.position[0].data(), hc ? b.color[0].data() : nullptr, hi ? b.intensity.data() : nullptr, nullptr, n, &n,
                                     nullptr));
        return b;
    }
    void write_to_directory(const std::string& dir) const { check(pcv_s2_write_dir(s_, dir.c_str())); }
    pcv_s2cloud* raw() const { return s_; }

   private:
    pcv_s2cloud* s_;
};

// S2Splitter::with_split_level(level, path, Encoding::Plain, ..) + write(batch)... + get_meta (read_write/s2.rs:33-50,59-125,165-173):
// the batches of one cloud, split by S2 cell on the GPU; throws ("... is not a valid ECEF point") like the writer's Err.
template <class BatchIterator>
inline S2Cells s2_split(Context& ctx, const std::string& output_directory, BatchIterator begin, BatchIterator end, uint32_t split_level = 20) {
    std::vector<std::array<double, 3>> pos;
    std::vector<std::array<uint8_t, 3>> col;
    std::vector<float> inten;
    for (BatchIterator it = begin; it != end; ++it) {
        const PointsBatch& b = *it;
        pos.insert(pos.end(), b.position.begin(), b.position.end());
        col.insert(col.end(), b.color.begin(), b.color.end());
        inten.insert(inten.end(), b.intensity.begin(), b.intensity.end());
    }
    pcv_points pts{};
    const double* base = pos.empty() ? nullptr : pos[0].data();
    pts.x = base;
    pts.y = base ? base + 1 : nullptr;
    pts.z = base ? base + 2 : nullptr;
    pts.stride = 3;
    pts.rgb = col.size() == pos.size() && !col.empty() ? col[0].data() : nullptr;
    pts.intensity = inten.size() == pos.size() && !inten.empty() ? inten.data() : nullptr;
    pts.n = pos.size();
    pcv_s2cloud* s = nullptr;
    check(pcv_s2_build(ctx.raw(), &pts, split_level, &s));
    S2Cells cloud(s);
    if (!output_directory.empty()) cloud.write_to_directory(output_directory);
    return cloud;
}

// ParallelIterator::new(point_clouds, query, batch_size, num_threads, buffer_size).try_for_each_batch(func) — iterator.rs:238-257.
// num_threads / buffer_size are accepted for signature parity; the GPU path streams on the caller's thread.
class ParallelIterator {
   public:
    ParallelIterator(const std::vector<const Octree*>& point_clouds, const PointQuery& query, size_t batch_size, size_t /*num_threads*/ = 1,
                     size_t /*buffer_size*/ = 4)
        : clouds_(point_clouds), query_(query), batch_size_(batch_size) {}
    // func returns true to continue, false to stop (== Err).  Returns true if every batch was consumed.
    bool try_for_each_batch(const std::function<bool(PointsBatch&&)>& func) {
        struct State {
            const std::function<bool(PointsBatch&&)>* f;
        } st{&func};
        auto tramp = [](void* user, const pcv_batch* b) -> int {
            State* s = (State*)user;
            PointsBatch pb;
            pb.position.resize(b->n);
            pb.color.resize(b->n);
            for (uint64_t i = 0; i < b->n; ++i) {
                pb.position[i] = {b->xyz[3 * i], b->xyz[3 * i + 1], b->xyz[3 * i + 2]};
                pb.color[i] = {b->rgb[3 * i], b->rgb[3 * i + 1], b->rgb[3 * i + 2]};
            }
            if (b->intensity) pb.intensity.assign(b->intensity, b->intensity + b->n);
            pb.source_index.assign(b->src_index, b->src_index + b->n);
            return (*s->f)(std::move(pb)) ? 0 : 1;
        };
        std::vector<pcv_interval> f;
        for (auto& iv : query_.filter_intervals) f.push_back(pcv_interval{iv.lower_bound, iv.upper_bound});
        for (const Octree* o : clouds_) {
            const int rc = pcv_query_points(o->raw(), &query_.location.raw, f.empty() ? nullptr : f.data(), (uint32_t)f.size(), batch_size_, tramp, &st);
            if (rc == PCV_ERR_CANCELLED) return false;
            check(rc);
        }
        return true;
    }

   private:
    std::vector<const Octree*> clouds_;
    PointQuery query_;
    size_t batch_size_;
};

}  // namespace pcv
