// build_host.hpp — host orchestration of the GPU octree build (backend agnostic).
//
// Replaces src/octree/generation.rs:289-403 (build_octree) with a different algorithm that yields the
// same tree, the same per-node point order and the same stored position codes:
//
//   split phase    the reference's recursive 8-way file split (generation.rs:58-193) becomes a
//                  top-down, level-synchronous stable multi-way partition: each pass resolves G octree
//                  levels for every point that still sits in a node with > MAX_POINTS_PER_NODE points
//                  (`hist` -> per-tile digit histograms, `scan` -> per-digit prefix over tiles, host
//                  decides leaf/split for the 8^1..8^G descendants, `scatter` -> stable partition
//                  into the next pass's segments or into the leaf arena).  The per-point digits come
//                  from the re-quantising descent in chain.h, so node membership is bit-identical.
//   subsample      the level-by-level rewrite (generation.rs:195-253,335-387) is replaced by its closed
//                  form: a point at rank j of a node X with parent P moves up iff j % 8 == 0, to rank
//                  off(X in P) + j/8; otherwise it stays at slot j - j/8 - 1.  `place` walks each leaf
//                  point up, re-encoding through every cube it passes (same DEC/ENC chain as the
//                  reference's rewrites), and writes it to its final node-contiguous slot.
//
// The Backend interface is implemented by the CUDA kernels (kernels_build.cuh).  A second, test-only
// implementation lives under tests/ to exercise this host logic without a GPU; the shipped library
// contains only the CUDA one.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "chain.h"

namespace pcv {

typedef unsigned __int128 u128;

struct PointsView {  // device pointers
    const double* x;
    const double* y;
    const double* z;
    uint64_t stride;
    const uint8_t* rgb;
    const float* intensity;
    uint64_t n;
};

// ---- structures shared with the kernels ---------------------------------------------------------
struct RecN {  // narrow partition record: codes of U8/U16/F32 levels
    uint32_t c[3];
    uint32_t idx;
};
struct RecW {  // wide record: used for the whole build when any level >= 1 is Float64 encoded
    uint64_t c[3];
    uint32_t idx;
    uint32_t pad;
};
struct TileDesc {
    uint64_t start;   // first record (or raw point) of the tile
    uint32_t count;   // <= kTilePoints
    uint32_t active;  // index into the pass's active-node array
};
struct ActiveDesc {
    double m[3];
    double e;
    uint64_t start, count;  // the node's segment in the pass's input (records, or raw points for the root)
    uint32_t chunk_begin, nchunks;
    uint32_t tile_begin;    // first tile of the node (tiles never straddle nodes)
    uint32_t pad0;
};
struct ChunkDesc {
    uint32_t tile_begin, ntiles;
    uint32_t active;
    uint32_t first;  // 1 if first chunk of its node
};
struct BucketDesc {
    uint64_t dest;    // first record of the bucket in its destination buffer
    uint16_t b0, b1;  // digit range [b0,b1) it collects (contiguous: a whole sub-tree)
    uint8_t keep;     // 1..G: which level's codes the destination stores
    uint8_t kind;     // 0 = next pass segment, 1 = leaf arena
    uint16_t pad;
};
struct DNode {
    double m[3];
    double e;
    double ry;                // RN(1/e)
    uint64_t off_in_parent;   // sum_{k'<k} ceil(n(P.k')/8)
    uint64_t out_point_off;
    uint64_t out_xyz_off;     // bytes
    uint64_t arena_off, count;  // leaves (and collectors in the top assembly): segment of the leaf arena
    int32_t parent;           // -1 for the root
    int32_t enc;
};
struct LeafTile {
    uint64_t arena_start;
    uint64_t j0;
    uint32_t node;
    uint32_t count;
};

constexpr uint32_t kTilePoints = 4096;   // partition tile (hist / scatter)
constexpr uint32_t kChunkTiles = 256;    // tiles per scan chunk
constexpr uint32_t kPlaceTile = 2048;    // place tile

struct PassArgs {
    int level, G, nbins;
    bool root, wide;
    PointsView pts;
    const void* rec_in;
    void* rec_next;
    void* arena;
    // colour travels with the records (packed r | g<<8 | b<<16) so the final placement does not gather it
    const uint32_t* col_in;
    uint32_t* col_next;
    uint32_t* col_arena;
    uint32_t ntiles, nactive, nchunks;
    uint64_t npoints;  // points still being partitioned in this pass
    const ActiveDesc* d_active;
    const ChunkDesc* d_chunks;
    uint32_t* d_tile_counts;  // [ntiles][nbins]; after scan: exclusive prefix over the node's tiles
    uint32_t* d_chunk_sums;   // [nchunks][nbins]
    uint64_t* d_node_bins;    // [nactive][nbins]
    const uint16_t* d_lut;    // [nactive][nbins] -> local bucket, 0xFFFF for empty digits
    const BucketDesc* d_buckets;  // [nactive][nbins]
    LevelTable lv;
};

struct PlaceArgs {
    bool wide;
    PointsView pts;
    const void* arena;
    const uint32_t* col_arena;
    int fast;  // LevelTable::fast
    const DNode* d_nodes;
    const uint32_t* d_leaf_tile_begin;  // [nleaves + 1] first tile of every leaf
    const uint32_t* d_leaf_node;        // [nleaves] index into d_nodes
    uint32_t nleaves;
    uint32_t ntiles;
    uint64_t npoints, xyz_bytes;
    uint32_t prefetch_tiles = 0;  // device backends: L2-prefetch the leaf tile this many blocks ahead (0 = off)
    uint8_t* out_xyz;
    uint8_t* out_rgb;
    float* out_intensity;
    uint32_t* out_src;
};

// Tile descriptors are not materialised: a block finds its node by binary search over the per-node first-tile index.
PCV_HD uint32_t upper_index(const uint32_t* begin, size_t stride_words, uint32_t n, uint32_t b) {
    uint32_t lo = 0, hi = n;  // largest i in [0, n) with begin[i] <= b
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (begin[(size_t)mid * stride_words] <= b)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}
PCV_HD TileDesc tile_of(const PassArgs& a, uint32_t b) {
    const uint32_t i = upper_index(&a.d_active[0].tile_begin, sizeof(ActiveDesc) / 4, a.nactive, b);
    const ActiveDesc& act = a.d_active[i];
    const uint64_t o = (uint64_t)(b - act.tile_begin) * kTilePoints;
    const uint64_t rem = act.count - o;
    return TileDesc{act.start + o, (uint32_t)(rem < kTilePoints ? rem : kTilePoints), i};
}
PCV_HD LeafTile leaf_tile_of(const PlaceArgs& a, uint32_t b) {
    const uint32_t l = upper_index(a.d_leaf_tile_begin, 1, a.nleaves, b);
    const uint32_t node = a.d_leaf_node[l];
    const DNode& nd = a.d_nodes[node];
    const uint64_t o = (uint64_t)(b - a.d_leaf_tile_begin[l]) * kPlaceTile;
    const uint64_t rem = nd.count - o;
    return LeafTile{nd.arena_off + o, o, node, (uint32_t)(rem < kPlaceTile ? rem : kPlaceTile)};
}

struct Backend {
    virtual ~Backend() {}
    virtual void* dmalloc(size_t bytes) = 0;
    virtual void dfree(void* p) = 0;
    virtual void h2d(void* d, const void* h, size_t bytes) = 0;
    virtual void d2h(void* h, const void* d, size_t bytes) = 0;
    virtual void hist(const PassArgs& a) = 0;
    virtual void scan(const PassArgs& a) = 0;
    virtual void scatter(const PassArgs& a) = 0;
    virtual void place(const PlaceArgs& a) = 0;
    virtual void mark(int what) {}  // timing hooks: 0 partition start, 1 partition end / place start, 2 place end
};

// ---- host-side node algebra (node.rs) -------------------------------------------------------------
struct HNode {
    u128 index;
    int level;
    int parent;
    int child[8];
    uint64_t count;  // points routed into the node by the split phase
    bool leaf;
    uint64_t arena_off;
    uint64_t n_sub;  // size when the node is subsampled into its parent
    uint64_t off_in_parent;
    uint64_t final_count;
    uint64_t out_point_off, out_xyz_off;
    double m[3];
    double e;
    int enc;
};

inline uint32_t rust_as_u32(double v) {
    if (!(v == v) || v <= 0.0) return 0u;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
// PositionEncoding::new (codec.rs:31-40)
inline int position_encoding_for(double edge, double resolution) {
    uint32_t min_bits = rust_as_u32(std::log2(edge / resolution)) + 1u;
    if (min_bits <= 8) return ENC_U8;
    if (min_bits <= 16) return ENC_U16;
    if (min_bits <= 24) return ENC_F32;
    return ENC_F64;
}

// `root_min`: min corner of the root cube.  When every coordinate inside the root cube has a magnitude in
// [2^-300, 2^300) on every axis, the numerators q - m of the whole tree are exactly 0 or in [2^-353, 2^500) (q and m are
// doubles of at least that magnitude, so a non-zero difference is at least half an ulp of 2^-300), i.e. inside
// div_known's proven range without looking at them: fast = 2 (the kernels then only check the raw inputs once).
inline LevelTable make_level_table(double root_edge, double resolution, const double* root_min = nullptr) {
    LevelTable t;
    double e = root_edge;
    t.last_level = kMaxLevels - 1;
    bool found = false;
    t.fast = 1;
    for (int L = 0; L < kMaxLevels; ++L) {
        t.edge[L] = e;
        t.ry[L] = 1.0 / e;  // RN(1/edge), for div_known
        if (!div_known_ok(e)) t.fast = 0;
        t.enc[L] = (int8_t)position_encoding_for(e, resolution);
        // should_split_node (generation.rs:128-150): a node at level L >= 1 is split only if edge > resolution.
        if (!found && L >= 1 && !(e > resolution)) {
            t.last_level = L;
            found = true;
        }
        e /= 2.;  // node.rs:161
    }
    if (t.fast && root_min && !std::getenv("PCV_CHECKED_FAST")) {  // PCV_CHECKED_FAST=1: diagnostic, keep the per-numerator checks
        const double lo = std::ldexp(1.0, -300), hi = std::ldexp(1.0, 300);
        bool ok = root_edge > 0.0 && root_edge < hi;
        for (int k = 0; k < 3 && ok; ++k) {
            const double a = root_min[k], b = root_min[k] + root_edge;
            ok = std::isfinite(a) && std::fabs(a) < hi && std::fabs(b) < hi && (a >= lo || b <= -lo);
        }
        if (ok) t.fast = 2;
    }
    return t;
}

struct BuildResult {
    std::vector<HNode> nodes;     // creation order (parents before children)
    std::vector<int> sorted;      // node indices sorted by NodeId
    uint64_t n = 0;
    uint64_t xyz_bytes = 0;
    uint8_t* d_xyz = nullptr;
    uint8_t* d_rgb = nullptr;
    float* d_intensity = nullptr;
    uint32_t* d_src = nullptr;
    uint32_t passes = 0;
    uint32_t deepest_level = 0;
    double host_ms_plan = 0, host_ms_wait = 0;
    uint64_t algorithmic_bytes = 0;
    LevelTable lv;
    double root_min[3];
    double root_edge;
};

struct BuildError : std::runtime_error {
    int code;
    BuildError(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

// Multi-GPU sharding: this context builds only the sub-trees below some level-k cells.  `counts` holds the
// GLOBAL point counts of every cell of levels 1..k (level j at offset (8^j - 8) / 7), so that nodes above level k take
// the same split decision on every rank; nodes of level k-1 ("collectors") keep the every-8th points of their local
// children (encoded in the collector's cube, in child order) for the top-of-tree assembly on one rank.
struct ShardSpec {
    int k = 0;
    const uint64_t* counts = nullptr;
    static size_t level_offset(int level) { return (((size_t)1 << (3 * level)) - 8) / 7; }
    uint64_t count_at(int level, uint64_t index) const { return counts[level_offset(level) + index]; }
};

class BuildPlan {
   public:
    Backend& be;
    uint64_t max_points;
    int G;
    ShardSpec shard;
    BuildPlan(Backend& b, uint64_t max_points_per_node, int levels_per_pass)
        : be(b), max_points(max_points_per_node ? max_points_per_node : 100000), G(levels_per_pass) {
        if (G < 1 || G > 3) G = 3;
    }

    template <class T>
    T* upload(const std::vector<T>& v, std::vector<void*>& owned) {
        if (v.empty()) return nullptr;
        T* d = (T*)be.dmalloc(v.size() * sizeof(T));
        owned.push_back(d);
        be.h2d(d, v.data(), v.size() * sizeof(T));
        return d;
    }

    BuildResult run(const PointsView& pts, double resolution, const double bmin[3], const double bmax[3]) {
        if (!(resolution > 0.0)) throw BuildError(-1, "resolution must be > 0");
        if (pts.n >= 0xFFFFFFFFull) throw BuildError(-6, "more than 2^32-2 points per context are not supported");
        BuildResult R;
        R.n = pts.n;
        // Cube::bounding (aabb.rs:149-157)
        double E = std::fmax(std::fmax(bmax[0] - bmin[0], bmax[1] - bmin[1]), bmax[2] - bmin[2]);
        for (int a = 0; a < 3; ++a) R.root_min[a] = bmin[a];
        R.root_edge = E;
        R.lv = make_level_table(E, resolution, bmin);
        const LevelTable& lv = R.lv;
        if (pts.n == 0) return R;  // no leaves -> no nodes at all (generation.rs:325-397)

        bool wide = false;
        for (int L = 1; L <= lv.last_level; ++L) wide = wide || lv.enc[L] == ENC_F64;
        const size_t rec_bytes = wide ? sizeof(RecW) : sizeof(RecN);

        std::vector<HNode>& nodes = R.nodes;
        {
            HNode r{};
            r.index = 0;
            r.level = 0;
            r.parent = -1;
            for (int k = 0; k < 8; ++k) r.child[k] = -1;
            r.count = pts.n;
            r.leaf = false;
            for (int a = 0; a < 3; ++a) r.m[a] = bmin[a];
            r.e = E;
            r.enc = lv.enc[0];
            nodes.push_back(r);
        }

        struct Active {
            int node;
            uint64_t start, count;
        };
        std::vector<Active> active{{0, 0, pts.n}};
        int L = 0;
        void* bufs[2] = {nullptr, nullptr};
        uint32_t* cols[2] = {nullptr, nullptr};
        void* arena = be.dmalloc((size_t)pts.n * rec_bytes);
        uint32_t* col_arena = (uint32_t*)be.dmalloc((size_t)pts.n * 4 + 64);  // + slack: the scatter's bulk copy reads whole 16-byte granules
        uint64_t arena_used = 0;
        int cur = -1;  // -1: raw input
        std::vector<void*> scratch;
        auto free_scratch = [&]() {
            for (void* p : scratch) be.dfree(p);
            scratch.clear();
        };
        be.mark(0);
        try {
            while (!active.empty()) {
                const auto tp0 = std::chrono::steady_clock::now();
                const int Gp = std::min(G, lv.last_level - L);
                if (Gp < 1) throw BuildError(-6, "octree deeper than 40 levels is not representable in NodeId");
                const int nbins = 1 << (3 * Gp);
                // ---- tiles / chunks ----
                uint32_t ntiles_total = 0;
                std::vector<ChunkDesc> chunks;
                std::vector<ActiveDesc> adesc(active.size());
                uint64_t active_points = 0;
                for (size_t a = 0; a < active.size(); ++a) {
                    const HNode& nd = nodes[active[a].node];
                    for (int k = 0; k < 3; ++k) adesc[a].m[k] = nd.m[k];
                    adesc[a].e = nd.e;
                    uint64_t c = active[a].count, s = active[a].start;
                    active_points += c;
                    const uint32_t t0 = ntiles_total;
                    const uint32_t nt = (uint32_t)((c + kTilePoints - 1) / kTilePoints);
                    ntiles_total += nt;
                    adesc[a].start = s;
                    adesc[a].count = c;
                    adesc[a].tile_begin = t0;
                    adesc[a].chunk_begin = (uint32_t)chunks.size();
                    adesc[a].nchunks = (nt + kChunkTiles - 1) / kChunkTiles;
                    for (uint32_t o = 0; o < nt; o += kChunkTiles)
                        chunks.push_back(ChunkDesc{t0 + o, std::min(kChunkTiles, nt - o), (uint32_t)a, o == 0 ? 1u : 0u});
                }
                PassArgs pa{};
                pa.level = L;
                pa.G = Gp;
                pa.nbins = nbins;
                pa.root = cur < 0;
                pa.wide = wide;
                pa.pts = pts;
                pa.rec_in = cur < 0 ? nullptr : bufs[cur];
                int nxt = cur < 0 ? 0 : 1 - cur;
                if (!bufs[nxt]) {
                    bufs[nxt] = be.dmalloc((size_t)pts.n * rec_bytes);
                    cols[nxt] = (uint32_t*)be.dmalloc((size_t)pts.n * 4 + 64);  // + slack: the scatter's bulk copy reads whole 16-byte granules
                }
                pa.rec_next = bufs[nxt];
                pa.arena = arena;
                pa.col_in = cur < 0 ? nullptr : cols[cur];
                pa.col_next = cols[nxt];
                pa.col_arena = col_arena;
                pa.ntiles = ntiles_total;
                pa.nactive = (uint32_t)active.size();
                pa.nchunks = (uint32_t)chunks.size();
                pa.npoints = active_points;
                pa.d_active = upload(adesc, scratch);
                pa.d_chunks = upload(chunks, scratch);
                pa.d_tile_counts = (uint32_t*)be.dmalloc((size_t)pa.ntiles * nbins * 4);
                scratch.push_back(pa.d_tile_counts);
                pa.d_chunk_sums = (uint32_t*)be.dmalloc((size_t)pa.nchunks * nbins * 4);
                scratch.push_back(pa.d_chunk_sums);
                pa.d_node_bins = (uint64_t*)be.dmalloc((size_t)pa.nactive * nbins * 8);
                scratch.push_back(pa.d_node_bins);
                pa.lv = lv;

                const auto tq0 = std::chrono::steady_clock::now();
                be.hist(pa);
                be.scan(pa);
                std::vector<uint64_t> bins((size_t)pa.nactive * nbins);
                const auto tw0 = std::chrono::steady_clock::now();
                R.host_ms_plan += std::chrono::duration<double, std::milli>(tw0 - tp0).count();
                be.d2h(bins.data(), pa.d_node_bins, bins.size() * 8);
                const auto tw1 = std::chrono::steady_clock::now();
                R.host_ms_wait += std::chrono::duration<double, std::milli>(tw1 - tw0).count();

                // ---- decide leaf / split for every descendant within Gp levels ----
                nodes.reserve(nodes.size() + (size_t)pa.nactive * 16 + 64);
                std::vector<uint16_t> lut((size_t)pa.nactive * nbins, 0xFFFF);
                std::vector<BucketDesc> buckets((size_t)pa.nactive * nbins);
                std::vector<Active> next_active;
                uint64_t next_used = 0;
                for (size_t a = 0; a < active.size(); ++a) {
                    const uint64_t* nb = &bins[a * nbins];
                    uint16_t nlocal = 0;
                    uint64_t total = 0;
                    for (int b = 0; b < nbins; ++b) total += nb[b];
                    if (total != active[a].count) throw BuildError(-2, "internal: histogram total mismatch");
                    // iterative expansion with an explicit stack: (node, sublevel j, b0, width)
                    struct Fr {
                        int node, j, b0, w;
                    };
                    Fr st[64];  // depth-first over <= 3 sub-levels: at most 8 pending frames per level
                    int sp = 0;
                    st[sp++] = Fr{active[a].node, 0, 0, nbins};
                    while (sp > 0) {
                        const Fr f = st[--sp];
                        int w = f.w / 8;
                        Fr pend[8];
                        int npend = 0;
                        for (int k = 0; k < 8; ++k) {
                            int b0 = f.b0 + k * w;
                            uint64_t cnt = 0;
                            for (int b = b0; b < b0 + w; ++b) cnt += nb[b];
                            if (cnt == 0) continue;
                            HNode c{};
                            const HNode& p = nodes[f.node];
                            c.level = p.level + 1;
                            c.index = (p.index << 3) + (u128)k;  // node.rs:120-125
                            c.parent = f.node;
                            for (int q = 0; q < 8; ++q) c.child[q] = -1;
                            c.count = cnt;
                            c.e = lv.edge[c.level];
                            // node.rs:165-170: x = bit2, y = bit1, z = bit0
                            c.m[0] = (k & 4) ? p.m[0] + c.e : p.m[0];
                            c.m[1] = (k & 2) ? p.m[1] + c.e : p.m[1];
                            c.m[2] = (k & 1) ? p.m[2] + c.e : p.m[2];
                            c.enc = lv.enc[c.level];
                            const uint64_t decision_cnt = (shard.k && c.level <= shard.k) ? shard.count_at(c.level, (uint64_t)c.index) : cnt;
                            bool split = decision_cnt > max_points && c.e > resolution;  // generation.rs:128-150
                            if (split && c.level >= kMaxLevels - 1)
                                throw BuildError(-6, "octree deeper than 40 levels is not representable in NodeId");
                            c.leaf = !split;
                            int ci = (int)nodes.size();
                            nodes.push_back(c);
                            nodes[f.node].child[k] = ci;
                            int j = f.j + 1;
                            if (split && j < Gp) {
                                pend[npend++] = Fr{ci, j, b0, w};
                                continue;
                            }
                            BucketDesc bd{};
                            bd.b0 = (uint16_t)b0;
                            bd.b1 = (uint16_t)(b0 + w);
                            bd.keep = (uint8_t)j;
                            if (split) {
                                bd.kind = 0;
                                bd.dest = next_used;
                                next_active.push_back(Active{ci, next_used, cnt});
                                next_used += cnt;
                            } else {
                                bd.kind = 1;
                                bd.dest = arena_used;
                                nodes[ci].arena_off = arena_used;
                                arena_used += cnt;
                                R.deepest_level = std::max<uint32_t>(R.deepest_level, (uint32_t)c.level);
                            }
                            for (int b = b0; b < b0 + w; ++b) lut[a * nbins + b] = nlocal;
                            buckets[a * nbins + nlocal] = bd;
                            ++nlocal;
                        }
                        for (int i = npend; i-- > 0;) st[sp++] = pend[i];
                    }
                }
                const auto tq1 = std::chrono::steady_clock::now();
                pa.d_lut = upload(lut, scratch);
                pa.d_buckets = upload(buckets, scratch);
                be.scatter(pa);
                if (std::getenv("PCV_TIMING"))
                    fprintf(stderr, "[pcv timing] pass L=%d G=%d active=%zu tiles=%u pts=%llu | descs %.2f wait %.2f decide %.2f upload+launch %.2f ms\n", L, Gp, active.size(),
                            pa.ntiles, (unsigned long long)active_points, std::chrono::duration<double, std::milli>(tq0 - tp0).count(),
                            std::chrono::duration<double, std::milli>(tw1 - tw0).count(), std::chrono::duration<double, std::milli>(tq1 - tw1).count(),
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tq1).count());
                free_scratch();
                R.host_ms_plan += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw1).count();
                R.passes++;
                active.swap(next_active);
                cur = nxt;
                L += Gp;
            }
        } catch (...) {
            free_scratch();
            be.dfree(arena);
            be.dfree(col_arena);
            for (void* b : bufs)
                if (b) be.dfree(b);
            for (uint32_t* b : cols)
                if (b) be.dfree(b);
            throw;
        }
        for (void* b : bufs)
            if (b) be.dfree(b);
        for (uint32_t* b : cols)
            if (b) be.dfree(b);
        be.mark(1);

        const bool dbg = std::getenv("PCV_TIMING") != nullptr;
        auto tnow = []() { return std::chrono::steady_clock::now(); };
        auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        const auto ts0 = tnow();
        // ---- subsample plan: closed form of generation.rs:195-253,335-387 ----
        for (size_t i = nodes.size(); i-- > 0;) {
            HNode& x = nodes[i];
            if (x.leaf) {
                x.n_sub = x.count;
            } else {
                uint64_t off = 0;
                for (int k = 0; k < 8; ++k) {
                    if (x.child[k] < 0) continue;
                    HNode& c = nodes[x.child[k]];
                    c.off_in_parent = off;
                    off += (c.n_sub + 7) / 8;  // every 8th point by current index: ceil(n/8)
                }
                x.n_sub = off;
                if (shard.k && x.level < shard.k - 1) x.n_sub = 0;  // above the collectors: assembled elsewhere
            }
        }
        auto is_collector = [&](const HNode& x) { return shard.k && x.level == shard.k - 1; };
        for (auto& x : nodes) x.final_count = (x.parent < 0 || is_collector(x)) ? x.n_sub : x.n_sub - (x.n_sub + 7) / 8;

        const auto ts1 = tnow();
        // ---- output layout: nodes sorted by NodeId (level << 120 | index) ----
        // (the device arrays are laid out in node creation order; the NodeId-sorted table is produced while the place
        // kernel runs)
        uint64_t poff = 0, boff = 0, algo_xyz = 0;
        for (size_t i = 0; i < nodes.size(); ++i) {
            HNode& x = nodes[i];
            x.out_point_off = poff;
            boff = (boff + 15) & ~15ull;  // node .xyz blocks are 16-byte aligned inside the device array
            x.out_xyz_off = boff;
            poff += x.final_count;
            boff += x.final_count * 3 * (uint64_t)enc_bytes(x.enc);
            algo_xyz += x.final_count * 3 * (uint64_t)enc_bytes(x.enc);
        }
        if (poff != pts.n) {
            be.dfree(arena);
            be.dfree(col_arena);
            throw BuildError(-2, "internal: subsample plan does not conserve points");
        }
        R.xyz_bytes = boff;
        R.algorithmic_bytes = 27ull * pts.n + algo_xyz + 3ull * pts.n + (pts.intensity ? 8ull * pts.n : 0ull);

        const auto ts2 = tnow();
        // ---- place ----
        std::vector<DNode> dn(nodes.size());
        std::vector<uint32_t> leaf_tile_begin, leaf_node;
        leaf_tile_begin.reserve(nodes.size() + 1);
        leaf_node.reserve(nodes.size());
        uint32_t nplace_tiles = 0;
        for (size_t i = 0; i < nodes.size(); ++i) {
            const HNode& x = nodes[i];
            DNode& d = dn[i];
            for (int a = 0; a < 3; ++a) d.m[a] = x.m[a];
            d.e = x.e;
            d.ry = 1.0 / x.e;
            d.off_in_parent = x.off_in_parent;
            d.out_point_off = x.out_point_off;
            d.out_xyz_off = x.out_xyz_off;
            d.arena_off = x.arena_off;
            d.count = x.leaf ? x.count : 0;
            d.parent = is_collector(x) ? -1 : x.parent;  // a collector ends the up-walk like the root does
            d.enc = x.enc;
            if (x.leaf) {
                leaf_tile_begin.push_back(nplace_tiles);
                leaf_node.push_back((uint32_t)i);
                nplace_tiles += (uint32_t)((x.count + kPlaceTile - 1) / kPlaceTile);
            }
        }
        leaf_tile_begin.push_back(nplace_tiles);
        const auto ts3 = tnow();
        R.d_xyz = (uint8_t*)be.dmalloc(std::max<uint64_t>(boff, 16));
        R.d_rgb = (uint8_t*)be.dmalloc((size_t)pts.n * 3);
        R.d_src = (uint32_t*)be.dmalloc((size_t)pts.n * 4 + 64);  // + slack: the scatter's bulk copy reads whole 16-byte granules
        R.d_intensity = pts.intensity ? (float*)be.dmalloc((size_t)pts.n * 4) : nullptr;
        const auto ts4 = tnow();
        PlaceArgs pl{};
        pl.wide = wide;
        pl.pts = pts;
        pl.arena = arena;
        pl.col_arena = col_arena;
        pl.fast = lv.fast;
        pl.d_nodes = upload(dn, scratch);
        pl.d_leaf_tile_begin = upload(leaf_tile_begin, scratch);
        pl.d_leaf_node = upload(leaf_node, scratch);
        pl.nleaves = (uint32_t)leaf_node.size();
        pl.ntiles = nplace_tiles;
        pl.npoints = pts.n;
        pl.xyz_bytes = algo_xyz;
        pl.out_xyz = R.d_xyz;
        pl.out_rgb = R.d_rgb;
        pl.out_intensity = R.d_intensity;
        pl.out_src = R.d_src;
        const auto ts5 = tnow();
        be.place(pl);
        be.mark(2);
        R.sorted.resize(nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i) R.sorted[i] = (int)i;
        std::sort(R.sorted.begin(), R.sorted.end(), [&](int a, int b) {
            if (nodes[a].level != nodes[b].level) return nodes[a].level < nodes[b].level;
            return nodes[a].index < nodes[b].index;
        });
        free_scratch();
        be.dfree(arena);
        be.dfree(col_arena);
        if (dbg)
            fprintf(stderr, "[pcv timing] plan %.2f  layout+sort %.2f  tables %.2f  alloc %.2f  upload %.2f  (nodes %zu, leaf tiles %zu)\n", tms(ts0, ts1), tms(ts1, ts2),
                    tms(ts2, ts3), tms(ts3, ts4), tms(ts4, ts5), nodes.size(), (size_t)nplace_tiles);
        return R;
    }
};

// Top-of-tree assembly for the sharded build: the nodes of levels 0..k-1, whose content is what their children gave up
// (every 8th point) - children being the level-k unit roots (built on their owner ranks) or other top nodes.
//   counts      global point counts of the cells of levels 1..k (ShardSpec layout)
//   unit_nsub   n(X) of every level-k cell when it is subsampled into its parent (0 for empty cells)
//   xyz/rgb/intensity: the collectors' gathered content (level k-1 nodes in index order, inside a node child order
//               0..7, inside a child rank order), positions as node-file bytes in the COLLECTOR's encoding.
// The place kernel then walks those points up exactly as in a single-GPU build.
inline BuildResult assemble_top(Backend& be, double resolution, const double bmin[3], const double bmax[3], int k, const uint64_t* counts,
                                const uint64_t* unit_nsub, const uint8_t* xyz, const uint8_t* rgb, const float* intensity, uint64_t npoints) {
    if (k < 1 || k > 3) throw BuildError(-1, "prefix levels must be 1..3");
    if (npoints >= 0xFFFFFFFFull) throw BuildError(-6, "too many points in the top assembly");
    BuildResult R;
    R.n = npoints;
    const double E = std::fmax(std::fmax(bmax[0] - bmin[0], bmax[1] - bmin[1]), bmax[2] - bmin[2]);
    for (int a = 0; a < 3; ++a) R.root_min[a] = bmin[a];
    R.root_edge = E;
    R.lv = make_level_table(E, resolution, bmin);
    const LevelTable& lv = R.lv;
    ShardSpec sp;
    sp.k = k;
    sp.counts = counts;
    uint64_t total = 0;
    for (uint64_t c = 0; c < 8; ++c) total += sp.count_at(1, c);
    if (total == 0) return R;
    bool wide = false;
    for (int L = 1; L <= lv.last_level; ++L) wide = wide || lv.enc[L] == ENC_F64;

    std::vector<HNode>& nodes = R.nodes;
    HNode r{};
    r.parent = -1;
    for (int q = 0; q < 8; ++q) r.child[q] = -1;
    for (int a = 0; a < 3; ++a) r.m[a] = bmin[a];
    r.e = E;
    r.enc = lv.enc[0];
    r.count = total;
    nodes.push_back(r);
    for (size_t i = 0; i < nodes.size(); ++i) {  // breadth first: levels 0..k-2 get their existing children
        if (nodes[i].level >= k - 1) continue;
        for (int c = 0; c < 8; ++c) {
            const HNode p = nodes[i];
            const uint64_t cidx = ((uint64_t)p.index << 3) + (uint64_t)c;
            const uint64_t cnt = sp.count_at(p.level + 1, cidx);
            if (cnt == 0) continue;
            HNode x{};
            x.level = p.level + 1;
            x.index = cidx;
            x.parent = (int)i;
            for (int q = 0; q < 8; ++q) x.child[q] = -1;
            x.count = cnt;
            x.e = lv.edge[x.level];
            x.m[0] = (c & 4) ? p.m[0] + x.e : p.m[0];
            x.m[1] = (c & 2) ? p.m[1] + x.e : p.m[1];
            x.m[2] = (c & 1) ? p.m[2] + x.e : p.m[2];
            x.enc = lv.enc[x.level];
            nodes[i].child[c] = (int)nodes.size();
            nodes.push_back(x);
        }
    }
    // n(X) bottom-up; collectors (level k-1) sum over the level-k unit roots
    std::vector<uint64_t> unit_off(nodes.size() * 8, 0);
    for (size_t i = nodes.size(); i-- > 0;) {
        HNode& x = nodes[i];
        uint64_t off = 0;
        if (x.level == k - 1) {
            for (int c = 0; c < 8; ++c) {
                unit_off[i * 8 + c] = off;
                off += (unit_nsub[((uint64_t)x.index << 3) + (uint64_t)c] + 7) / 8;
            }
        } else {
            for (int c = 0; c < 8; ++c) {
                if (x.child[c] < 0) continue;
                nodes[x.child[c]].off_in_parent = off;
                off += (nodes[x.child[c]].n_sub + 7) / 8;
            }
        }
        x.n_sub = off;
    }
    for (auto& x : nodes) x.final_count = x.parent < 0 ? x.n_sub : x.n_sub - (x.n_sub + 7) / 8;
    R.sorted.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) R.sorted[i] = (int)i;
    std::sort(R.sorted.begin(), R.sorted.end(), [&](int a, int b) {
        if (nodes[a].level != nodes[b].level) return nodes[a].level < nodes[b].level;
        return nodes[a].index < nodes[b].index;
    });
    uint64_t poff = 0, boff = 0, arena_total = 0;
    for (int i : R.sorted) {
        HNode& x = nodes[i];
        x.out_point_off = poff;
        boff = (boff + 15) & ~15ull;
        x.out_xyz_off = boff;
        poff += x.final_count;
        boff += x.final_count * 3 * (uint64_t)enc_bytes(x.enc);
        if (x.level == k - 1) {
            x.arena_off = arena_total;
            arena_total += x.n_sub;
        }
    }
    if (arena_total != npoints || poff != npoints) throw BuildError(-1, "top assembly: gathered point count does not match the unit sizes");
    R.xyz_bytes = boff;

    // records from the gathered node-file bytes (collector encoding), in arena order == gathered order
    const int cenc = lv.enc[k - 1], bpc = enc_bytes(cenc);
    const size_t rec_bytes = wide ? sizeof(RecW) : sizeof(RecN);
    std::vector<uint8_t> recs((size_t)npoints * rec_bytes);
    std::vector<uint32_t> cols(npoints);
    for (uint64_t i = 0; i < npoints; ++i) {
        uint64_t c[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) std::memcpy(&c[a], xyz + (i * 3 + a) * bpc, (size_t)bpc);
        if (wide) {
            RecW w{{c[0], c[1], c[2]}, (uint32_t)i, 0};
            std::memcpy(&recs[i * rec_bytes], &w, sizeof w);
        } else {
            RecN w{{(uint32_t)c[0], (uint32_t)c[1], (uint32_t)c[2]}, (uint32_t)i};
            std::memcpy(&recs[i * rec_bytes], &w, sizeof w);
        }
        cols[i] = (uint32_t)rgb[3 * i] | ((uint32_t)rgb[3 * i + 1] << 8) | ((uint32_t)rgb[3 * i + 2] << 16);
    }
    std::vector<void*> scratch;
    auto up = [&](const void* h, size_t bytes) {
        void* d = be.dmalloc(bytes ? bytes : 16);
        scratch.push_back(d);
        if (bytes) be.h2d(d, h, bytes);
        return d;
    };
    std::vector<DNode> dn(nodes.size());
    std::vector<uint32_t> leaf_tile_begin, leaf_node;
    uint32_t nplace_tiles = 0;
    for (size_t i = 0; i < nodes.size(); ++i) {
        const HNode& x = nodes[i];
        DNode& d = dn[i];
        for (int a = 0; a < 3; ++a) d.m[a] = x.m[a];
        d.e = x.e;
        d.ry = 1.0 / x.e;
        d.off_in_parent = x.off_in_parent;
        d.out_point_off = x.out_point_off;
        d.out_xyz_off = x.out_xyz_off;
        d.arena_off = x.arena_off;
        d.count = x.level == k - 1 ? x.n_sub : 0;
        d.parent = x.parent;
        d.enc = x.enc;
        if (x.level == k - 1 && x.n_sub) {
            leaf_tile_begin.push_back(nplace_tiles);
            leaf_node.push_back((uint32_t)i);
            nplace_tiles += (uint32_t)((x.n_sub + kPlaceTile - 1) / kPlaceTile);
        }
    }
    leaf_tile_begin.push_back(nplace_tiles);
    PlaceArgs pl{};
    pl.wide = wide;
    pl.pts = PointsView{nullptr, nullptr, nullptr, 1, nullptr, intensity ? (const float*)up(intensity, (size_t)npoints * 4) : nullptr, npoints};
    pl.arena = up(recs.data(), recs.size());
    pl.col_arena = (const uint32_t*)up(cols.data(), cols.size() * 4);
    pl.fast = lv.fast;
    pl.d_nodes = (const DNode*)up(dn.data(), dn.size() * sizeof(DNode));
    pl.d_leaf_tile_begin = (const uint32_t*)up(leaf_tile_begin.data(), leaf_tile_begin.size() * 4);
    pl.d_leaf_node = (const uint32_t*)up(leaf_node.data(), leaf_node.size() * 4);
    pl.nleaves = (uint32_t)leaf_node.size();
    pl.ntiles = nplace_tiles;
    pl.npoints = npoints;
    pl.xyz_bytes = boff;
    R.d_xyz = (uint8_t*)be.dmalloc(std::max<uint64_t>(boff, 16));
    R.d_rgb = (uint8_t*)be.dmalloc(std::max<uint64_t>(npoints * 3, 16));
    R.d_src = (uint32_t*)be.dmalloc(std::max<uint64_t>(npoints * 4, 16));
    R.d_intensity = intensity ? (float*)be.dmalloc(std::max<uint64_t>(npoints * 4, 16)) : nullptr;
    pl.out_xyz = R.d_xyz;
    pl.out_rgb = R.d_rgb;
    pl.out_intensity = R.d_intensity;
    pl.out_src = R.d_src;
    be.place(pl);
    for (void* p : scratch) be.dfree(p);
    return R;
}

}  // namespace pcv
