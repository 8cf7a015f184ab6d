// query_api.inl — query-side entry points of the C ABI (included at the end of pcv_api.cu).

namespace {

using namespace pcv;

struct Scratch {  // stream-ordered device allocations released on scope exit
    pcv_ctx* c;
    std::vector<void*> ptrs;
    explicit Scratch(pcv_ctx* ctx) : c(ctx) {}
    template <class T>
    T* alloc(size_t n) {
        T* p = (T*)c->be->dmalloc(n * sizeof(T));
        ptrs.push_back(p);
        return p;
    }
    template <class T>
    T* upload(const T* h, size_t n) {
        T* p = alloc<T>(n ? n : 1);
        if (n) c->be->h2d(p, h, n * sizeof(T));
        return p;
    }
    ~Scratch() {
        for (void* p : ptrs) c->be->dfree(p);
    }
};

struct QTables {
    std::vector<QNode> qn;
    std::vector<uint32_t> level_start;  // nlevels + 1
};

// Node table for the query kernels + parent/children indices (nodes are sorted by NodeId = level-major).
void ensure_tables(pcv_octree* o) {
    if (o->tables_ready) return;
    const size_t n = o->nodes.size();
    std::vector<QNode> qn(n);
    o->parent_of.assign(n, -1);
    o->children_of.assign(n * 8, -1);
    for (size_t i = 0; i < n; ++i) {
        const pcv_node_meta& m = o->nodes[i];
        QNode& q = qn[i];
        for (int a = 0; a < 3; ++a) q.m[a] = m.cube_min[a];
        q.e = m.cube_edge;
        q.point_off = m.point_offset;
        q.xyz_off = m.xyz_byte_offset;
        q.n = (uint32_t)m.num_points;
        q.enc = m.position_encoding;
        q.level = m.level;
        q.parent = -1;
        if (m.level > 0) {
            const u128 id = ((u128)m.id_high << 64) | m.id_low;
            const u128 idx = id & ((((u128)1) << 120) - 1);
            const u128 pid = ((u128)(m.level - 1) << 120) | (idx >> 3);  // node.rs:136-144
            const int p = o->find((uint64_t)(pid >> 64), (uint64_t)pid);
            q.parent = p;
            o->parent_of[i] = p;
            if (p >= 0) o->children_of[(size_t)p * 8 + (size_t)(idx & 7)] = (int32_t)i;
        }
    }
    pcv_ctx* c = o->ctx;
    if (n) {
        o->d_qnodes = c->be->dmalloc(n * sizeof(QNode));
        c->be->h2d(o->d_qnodes, qn.data(), n * sizeof(QNode));
        o->d_children = (int32_t*)c->be->dmalloc(n * 8 * sizeof(int32_t));
        c->be->h2d(o->d_children, o->children_of.data(), n * 8 * sizeof(int32_t));
    }
    o->tables_ready = true;
}

std::vector<uint32_t> level_starts(const pcv_octree* o) {
    int maxl = 0;
    for (const auto& m : o->nodes) maxl = std::max(maxl, m.level);
    std::vector<uint32_t> ls((size_t)maxl + 2, 0);
    for (const auto& m : o->nodes) ls[(size_t)m.level + 1]++;
    for (size_t i = 1; i < ls.size(); ++i) ls[i] += ls[i - 1];
    return ls;
}

// pass[loc][node] for nloc locations (BFS semantics), left on the device.
uint8_t* run_sat_device(pcv_octree* o, const std::vector<QueryGeom>& geoms, Scratch& s, const QueryGeom** d_geoms_out) {
    pcv_ctx* c = o->ctx;
    const uint32_t nn = (uint32_t)o->nodes.size(), nloc = (uint32_t)geoms.size();
    const QueryGeom* dg = s.upload(geoms.data(), geoms.size());
    if (d_geoms_out) *d_geoms_out = dg;
    if (nn == 0 || nloc == 0) return nullptr;
    std::vector<uint32_t> ls = level_starts(o);
    const uint32_t* dls = s.upload(ls.data(), ls.size());
    uint8_t* drel = s.alloc<uint8_t>((size_t)nn * nloc);
    uint8_t* dpass = s.alloc<uint8_t>((size_t)nn * nloc);
    dim3 grid((nn + 255) / 256, nloc);
    k_sat_nodes<<<grid, 256, 0, c->stream>>>(dg, (const QNode*)o->d_qnodes, nn, drel);
    k_propagate<<<nloc, 1024, 0, c->stream>>>((const QNode*)o->d_qnodes, dls, (int)ls.size() - 1, nn, drel, dpass);
    c->be->launches += 2;
    CU(cudaGetLastError());
    return dpass;
}

// pass[loc][node] for nloc locations (BFS semantics).  Returns host matrix.
std::vector<uint8_t> run_sat(pcv_octree* o, const std::vector<QueryGeom>& geoms, Scratch& s, const QueryGeom** d_geoms_out) {
    pcv_ctx* c = o->ctx;
    const uint32_t nn = (uint32_t)o->nodes.size(), nloc = (uint32_t)geoms.size();
    const QueryGeom* dg = s.upload(geoms.data(), geoms.size());
    if (d_geoms_out) *d_geoms_out = dg;
    std::vector<uint8_t> pass((size_t)nn * nloc);
    if (nn == 0 || nloc == 0) return pass;
    std::vector<uint32_t> ls = level_starts(o);
    const uint32_t* dls = s.upload(ls.data(), ls.size());
    uint8_t* drel = s.alloc<uint8_t>((size_t)nn * nloc);
    uint8_t* dpass = s.alloc<uint8_t>((size_t)nn * nloc);
    dim3 grid((nn + 255) / 256, nloc);
    k_sat_nodes<<<grid, 256, 0, c->stream>>>(dg, (const QNode*)o->d_qnodes, nn, drel);
    k_propagate<<<nloc, 1024, 0, c->stream>>>((const QNode*)o->d_qnodes, dls, (int)ls.size() - 1, nn, drel, dpass);
    c->be->launches += 2;
    CU(cudaGetLastError());
    c->be->d2h(pass.data(), dpass, pass.size());
    return pass;
}

// Rust std::collections::BinaryHeap<OpenNode> (max-heap on size_on_screen) restated: push = append + sift_up;
// pop = take last, swap into the root, sift the hole down to the bottom preferring the right child when
// left <= right, then sift_up (octree/mod.rs:360-404).
struct Open {
    int node;
    uint8_t rel;
    double size;
};
struct OpenHeap {
    std::vector<Open> v;
    void up(size_t pos) {
        const Open x = v[pos];
        while (pos > 0) {
            const size_t par = (pos - 1) >> 1;
            if (x.size <= v[par].size) break;
            v[pos] = v[par];
            pos = par;
        }
        v[pos] = x;
    }
    void push(const Open& x) {
        v.push_back(x);
        up(v.size() - 1);
    }
    bool pop(Open& out) {
        if (v.empty()) return false;
        Open last = v.back();
        v.pop_back();
        if (v.empty()) {
            out = last;
            return true;
        }
        out = v[0];
        const size_t end = v.size();
        size_t pos = 0, child = 1;
        while (child < end) {
            const size_t right = child + 1;
            if (right < end && !(v[child].size > v[right].size)) child = right;
            v[pos] = v[child];
            pos = child;
            child = 2 * pos + 1;
        }
        v[pos] = last;
        up(pos);
        return true;
    }
};

struct HostBatchBuf {  // PointStream (iterator.rs:123-166): re-chunk into exactly batch_size points
    std::vector<double> xyz;
    std::vector<uint8_t> rgb;
    std::vector<float> inten;
    std::vector<uint64_t> src;
    size_t head = 0;  // points already delivered from the front
    size_t size() const { return src.size() - head; }
    int deliver(size_t n, bool has_i, pcv_batch_cb cb, void* user) {
        pcv_batch b;
        b.n = n;
        b.xyz = xyz.data() + 3 * head;
        b.rgb = rgb.data() + 3 * head;
        b.intensity = has_i ? inten.data() + head : nullptr;
        b.src_index = src.data() + head;
        head += n;
        return cb(user, &b);
    }
    void compact(bool has_i) {
        if (head == 0) return;
        xyz.erase(xyz.begin(), xyz.begin() + 3 * head);
        rgb.erase(rgb.begin(), rgb.begin() + 3 * head);
        if (has_i) inten.erase(inten.begin(), inten.begin() + head);
        src.erase(src.begin(), src.begin() + head);
        head = 0;
    }
};

void make_tiles(const pcv_octree* o, uint32_t loc, uint32_t node, std::vector<QTile>& tiles) {
    const uint32_t n = (uint32_t)o->nodes[node].num_points;
    for (uint32_t f = 0; f < n; f += kQueryTile) tiles.push_back(QTile{loc, node, f, std::min(kQueryTile, n - f)});
}

struct CullResult {
    uint64_t total = 0;
    double* d_xyz = nullptr;
    uint8_t* d_rgb = nullptr;
    float* d_inten = nullptr;
    uint32_t* d_src = nullptr;
};

// count -> scan -> write for a list of tiles (host list, or a device list when d_tiles != nullptr).
CullResult run_cull(pcv_octree* o, const QueryGeom* d_geoms, const std::vector<QTile>& tiles, const pcv_interval* filters, uint32_t nfilt,
                    Scratch& s, unsigned long long* d_kept, unsigned long long* d_tested, const QTile* d_tiles = nullptr, uint64_t n_dtiles = 0) {
    pcv_ctx* c = o->ctx;
    CullResult r;
    if (tiles.empty() && n_dtiles == 0) return r;
    CullArgs a{};
    a.geoms = d_geoms;
    a.nodes = (const QNode*)o->d_qnodes;
    a.tiles = d_tiles ? d_tiles : s.upload(tiles.data(), tiles.size());
    a.xyz = o->d_xyz;
    a.rgb = o->d_rgb;
    a.intensity = o->d_intensity;
    a.src = o->d_src;
    a.filters = nfilt ? s.upload(filters, nfilt) : nullptr;
    a.nfilt = nfilt;
    const uint32_t nt = d_tiles ? (uint32_t)n_dtiles : (uint32_t)tiles.size();
    a.tile_keep = s.alloc<uint32_t>(nt);
    k_cull<false><<<nt, 256, 0, c->stream>>>(a);
    if (d_kept) k_tile_totals<<<(nt + 255) / 256, 256, 0, c->stream>>>(a.tiles, a.tile_keep, nt, d_kept, d_tested);
    unsigned long long* d_total = s.alloc<unsigned long long>(1);
    k_scan_u32<<<1, 1024, 0, c->stream>>>(a.tile_keep, nt, d_total);
    c->be->launches += d_kept ? 3 : 2;
    CU(cudaGetLastError());
    unsigned long long total = 0;
    c->be->d2h(&total, d_total, 8);
    r.total = total;
    if (total >= 0xFFFFFFFFull) throw BuildError(PCV_ERR_UNSUPPORTED, "more than 2^32-1 survivors in one launch");
    if (total == 0) return r;
    r.d_xyz = a.out_xyz = s.alloc<double>(3 * total);
    r.d_rgb = a.out_rgb = s.alloc<uint8_t>(3 * total);
    r.d_inten = a.out_intensity = o->d_intensity ? s.alloc<float>(total) : nullptr;
    r.d_src = a.out_src = s.alloc<uint32_t>(total);
    k_cull<true><<<nt, 256, 0, c->stream>>>(a);
    c->be->launches += 1;
    CU(cudaGetLastError());
    return r;
}

}  // namespace

extern "C" {

int pcv_nodes_in_location(const pcv_octree* oc, const pcv_location* loc, uint64_t* ids, uint64_t cap, uint64_t* n_out) {
    if (!oc || !loc || !n_out) return fail(PCV_ERR_INVALID, "null argument");
    if (loc->kind < 0 || loc->kind > 3) return fail(PCV_ERR_INVALID, "unknown location kind %d", loc->kind);
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ensure_tables(o);
    Scratch s(c);
    std::vector<QueryGeom> geoms{make_query_geom(*loc)};
    std::vector<uint8_t> pass = run_sat(o, geoms, s, nullptr);
    uint64_t n = 0;
    for (size_t i = 0; i < pass.size(); ++i)
        if (pass[i]) {
            if (n < cap && ids) {
                ids[2 * n] = o->nodes[i].id_high;
                ids[2 * n + 1] = o->nodes[i].id_low;
            }
            ++n;
        }
    *n_out = n;
    if (n > cap) return fail(PCV_ERR_INVALID, "capacity %llu < %llu nodes", (unsigned long long)cap, (unsigned long long)n);
    return PCV_OK;
    API_CATCH
}

int pcv_visible_nodes(const pcv_octree* oc, const double M[16], uint64_t* ids, uint64_t cap, uint64_t* n_out) {
    if (!oc || !M || !n_out) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ensure_tables(o);
    pcv_location loc{};
    loc.kind = PCV_LOC_FRUSTUM;
    memcpy(loc.clip_from_query, M, sizeof(double) * 16);
    if (!mat4_try_inverse(M, loc.query_from_clip)) return fail(PCV_ERR_SINGULAR, "Invalid projection matrix.");
    *n_out = 0;
    const uint32_t nn = (uint32_t)o->nodes.size();
    if (nn == 0) return PCV_OK;
    Scratch s(c);
    QueryGeom geom = make_query_geom(loc);
    const QueryGeom* dg = s.upload(&geom, 1);
    const double* dM = s.upload(M, 16);
    uint8_t* drel = s.alloc<uint8_t>(nn);
    double* dsize = s.alloc<double>(nn);
    k_visible_eval<<<(nn + 255) / 256, 256, 0, c->stream>>>(dg, dM, (const QNode*)o->d_qnodes, nn, drel, dsize);
    c->be->launches++;
    CU(cudaGetLastError());
    std::vector<uint8_t> rel(nn);
    std::vector<double> size(nn);
    c->be->d2h(rel.data(), drel, nn);
    c->be->d2h(size.data(), dsize, (size_t)nn * 8);
    // best-first traversal (octree/mod.rs:232-283).  The reference projects a node's corners when it pushes the node
    // (maybe_push_node -> relative_size_on_screen), so a corner with w == 0 (bit 7 of rel) is an error only for nodes that
    // are actually pushed - children of Out nodes are never looked at.
    const uint8_t kBadW = 0x80;
    auto bad_w = [&](int node) { return fail(PCV_ERR_INVALID, "projection of a corner of node %s has w == 0 (the reference panics here)",
                                             node_name(o->nodes[node].id_high, o->nodes[node].id_low).c_str()); };
    OpenHeap open;
    const int root = o->find(0, 0);
    if (root >= 0) {
        if (rel[root] & kBadW) return bad_w(root);
        open.push(Open{root, REL_CROSS, size[root]});
    }
    uint64_t n = 0;
    Open cur;
    while (open.pop(cur)) {
        for (int k = 0; k < 8; ++k) {
            const int ch = o->children_of[(size_t)cur.node * 8 + k];
            if (ch < 0) continue;  // maybe_push_node: only ids present in the meta
            const uint8_t r = rel[ch] & 0x7f;
            if (cur.rel == REL_CROSS) {
                if (r == REL_OUT) continue;
                if (rel[ch] & kBadW) return bad_w(ch);
                open.push(Open{ch, r, size[ch]});
            } else {
                if (rel[ch] & kBadW) return bad_w(ch);
                open.push(Open{ch, REL_IN, size[ch]});
            }
        }
        if (o->nodes[cur.node].num_points != 0) {
            if (n < cap && ids) {
                ids[2 * n] = o->nodes[cur.node].id_high;
                ids[2 * n + 1] = o->nodes[cur.node].id_low;
            }
            ++n;
        }
    }
    *n_out = n;
    if (n > cap) return fail(PCV_ERR_INVALID, "capacity %llu < %llu nodes", (unsigned long long)cap, (unsigned long long)n);
    return PCV_OK;
    API_CATCH
}

int pcv_query_points(const pcv_octree* oc, const pcv_location* loc, const pcv_interval* filters, uint32_t nfilt, uint64_t batch_size,
                     pcv_batch_cb cb, void* user) {
    if (!oc || !loc || !cb || batch_size == 0) return fail(PCV_ERR_INVALID, "null argument or batch_size == 0");
    if (nfilt && !filters) return fail(PCV_ERR_INVALID, "filters is null");
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    if (nfilt && !o->has_intensity) return fail(PCV_ERR_INVALID, "Filter attribute needs to be specified as query attribute.");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ensure_tables(o);
    std::vector<QueryGeom> geoms{make_query_geom(*loc)};
    std::vector<uint32_t> visit;
    const QueryGeom* dg = nullptr;
    Scratch sg(c);
    {
        std::vector<uint8_t> pass = run_sat(o, geoms, sg, &dg);
        for (size_t i = 0; i < pass.size(); ++i)
            if (pass[i] && o->nodes[i].num_points > 0) visit.push_back((uint32_t)i);
    }
    const bool has_i = o->has_intensity;
    HostBatchBuf hb;
    const uint64_t kMaxTested = 32ull << 20;  // points decoded per launch group
    size_t vi = 0;
    while (vi < visit.size()) {
        std::vector<QTile> tiles;
        uint64_t tested = 0;
        while (vi < visit.size() && (tiles.empty() || tested + (uint64_t)o->nodes[visit[vi]].num_points <= kMaxTested)) {
            make_tiles(o, 0, visit[vi], tiles);
            tested += (uint64_t)o->nodes[visit[vi]].num_points;
            ++vi;
        }
        Scratch s(c);
        CullResult r = run_cull(o, dg, tiles, filters, nfilt, s, nullptr, nullptr);
        if (r.total == 0) continue;
        hb.compact(has_i);
        const size_t old = hb.src.size(), tot = (size_t)r.total;
        hb.xyz.resize(3 * (old + tot));
        hb.rgb.resize(3 * (old + tot));
        if (has_i) hb.inten.resize(old + tot);
        hb.src.resize(old + tot);
        std::vector<uint32_t> src32(tot);
        CU(cudaMemcpyAsync(hb.xyz.data() + 3 * old, r.d_xyz, tot * 24, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(hb.rgb.data() + 3 * old, r.d_rgb, tot * 3, cudaMemcpyDeviceToHost, c->stream));
        if (has_i) CU(cudaMemcpyAsync(hb.inten.data() + old, r.d_inten, tot * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(src32.data(), r.d_src, tot * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < tot; ++i) hb.src[old + i] = src32[i];
        while (hb.size() >= batch_size)  // push_points_and_callback, iterator.rs:159-165
            if (hb.deliver((size_t)batch_size, has_i, cb, user)) return fail(PCV_ERR_CANCELLED, "cancelled by the consumer callback");
    }
    if (hb.size() > 0)  // last (short) batch, iterator.rs:316-323
        if (hb.deliver(hb.size(), has_i, cb, user)) return fail(PCV_ERR_CANCELLED, "cancelled by the consumer callback");
    return PCV_OK;
    API_CATCH
}

int pcv_query_batch_device(const pcv_octree* oc, const pcv_location* locs, uint32_t nloc, const pcv_interval* filters, uint32_t nfilt,
                           uint64_t* counts_out, uint64_t* tested_out) {
    if (!oc || (!locs && nloc)) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    if (nfilt && !o->has_intensity) return fail(PCV_ERR_INVALID, "Filter attribute needs to be specified as query attribute.");
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ensure_tables(o);
    c->qstats = pcv_query_stats{};
    if (nloc == 0) return PCV_OK;
    std::vector<QueryGeom> geoms(nloc);
    for (uint32_t i = 0; i < nloc; ++i) {
        if (locs[i].kind < 0 || locs[i].kind > 3) return fail(PCV_ERR_INVALID, "unknown location kind %d", locs[i].kind);
        geoms[i] = make_query_geom(locs[i]);  // per-location axis caching (sat.rs:111-143), host side: O(1) per location
    }
    const uint32_t nn = (uint32_t)o->nodes.size();
    for (uint32_t i = 0; i < nloc; ++i) {
        if (counts_out) counts_out[i] = 0;
        if (tested_out) tested_out[i] = 0;
    }
    if (nn == 0) return PCV_OK;
    const uint64_t l0 = c->be->launches;
    Scratch s(c);
    const QueryGeom* dg = s.upload(geoms.data(), geoms.size());
    cudaEvent_t ev[4];
    for (auto& e : ev) CU(cudaEventCreate(&e));
    CU(cudaEventRecord(ev[0], c->stream));
    // ---- node selection: level-synchronous frontier over all locations ----
    LocProj* dproj = s.alloc<LocProj>(nloc);
    k_loc_proj<<<nloc, 32, 0, c->stream>>>(dg, dproj);
    int maxl = 0;
    for (const auto& m : o->nodes) maxl = std::max(maxl, m.level);
    const uint32_t cap = (uint32_t)std::min<uint64_t>((uint64_t)nloc * nn, 48ull << 20);
    uint2* fr[2] = {s.alloc<uint2>(cap), s.alloc<uint2>(cap)};
    uint2* dpairs = s.alloc<uint2>(cap);
    // counters: [0 .. maxl + 1] frontier sizes, then npairs (u32); ntiles, bytes, cursor, out cursor (u64); overflow
    uint32_t* dcnt32 = s.alloc<uint32_t>((size_t)maxl + 4);
    unsigned long long* dcnt64 = s.alloc<unsigned long long>(4);
    int* dover = s.alloc<int>(1);
    unsigned long long* dk = s.alloc<unsigned long long>(nloc);
    unsigned long long* dt = s.alloc<unsigned long long>(nloc);
    c->be->zero(dcnt32, ((size_t)maxl + 4) * 4);
    c->be->zero(dcnt64, 4 * 8);
    c->be->zero(dover, 4);
    c->be->zero(dk, (size_t)nloc * 8);
    c->be->zero(dt, (size_t)nloc * 8);
    const int root = o->find(0, 0);
    if (root < 0) return fail(PCV_ERR_INVALID, "octree without a root node");
    k_bfs_seed<<<(nloc + 255) / 256, 256, 0, c->stream>>>(fr[0], nloc, (uint32_t)root, dcnt32);
    uint32_t* dnpairs = dcnt32 + maxl + 2;
    for (int L = 0; L <= maxl; ++L) {
        BfsArgs b{};
        b.geoms = dg;
        b.proj = dproj;
        b.nodes = (const QNode*)o->d_qnodes;
        b.children = o->d_children;
        b.fin = fr[L & 1];
        b.fout = fr[(L + 1) & 1];
        b.nin = dcnt32 + L;
        b.nout = dcnt32 + L + 1;
        b.cap = cap;
        b.pairs = dpairs;
        b.npairs = dnpairs;
        b.ntiles = dcnt64;
        b.tested = dt;
        b.bytes = dcnt64 + 1;
        b.overflow = dover;
        k_bfs_level<<<c->sm_count * 4, 256, 0, c->stream>>>(b);
    }
    c->be->launches += 3 + (uint64_t)maxl;
    CU(cudaGetLastError());
    unsigned long long h64[4];
    uint32_t npairs = 0;
    int over = 0;
    c->be->d2h(h64, dcnt64, sizeof h64);
    c->be->d2h(&npairs, dnpairs, 4);
    c->be->d2h(&over, dover, 4);
    if (over) return fail(PCV_ERR_UNSUPPORTED, "batched query visits more than %u (location, node) pairs per level; split the batch", cap);
    const unsigned long long ntl = h64[0];
    if (ntl >= 0xFFFFFFFFull) return fail(PCV_ERR_UNSUPPORTED, "work list too large (%llu tiles); split the batch", ntl);
    std::vector<unsigned long long> hk(nloc, 0), ht(nloc, 0);
    unsigned long long stored = 0;
    if (ntl) {
        QTile* dtiles = s.alloc<QTile>(ntl);
        k_pairs_to_tiles<<<c->sm_count * 4, 256, 0, c->stream>>>(dpairs, dnpairs, cap, (const QNode*)o->d_qnodes, dcnt64 + 2, dtiles);
        CU(cudaEventRecord(ev[1], c->stream));
        // ---- single-pass cull: survivors compacted into one output set (capacity bounded; the rest is counted only) ----
        unsigned long long tested_total = 0;
        c->be->d2h(ht.data(), dt, (size_t)nloc * 8);
        for (auto v : ht) tested_total += v;
        const unsigned long long outcap = std::min<unsigned long long>(tested_total, 192ull << 20);
        CullFusedArgs f{};
        f.c.geoms = dg;
        f.c.nodes = (const QNode*)o->d_qnodes;
        f.c.tiles = dtiles;
        f.c.xyz = o->d_xyz;
        f.c.rgb = o->d_rgb;
        f.c.intensity = o->d_intensity;
        f.c.src = o->d_src;
        f.c.filters = nfilt ? s.upload(filters, nfilt) : nullptr;
        f.c.nfilt = nfilt;
        f.c.out_xyz = s.alloc<double>(3 * outcap + 1);
        f.c.out_rgb = s.alloc<uint8_t>(3 * outcap + 1);
        f.c.out_intensity = o->d_intensity ? s.alloc<float>(outcap + 1) : nullptr;
        f.c.out_src = s.alloc<uint32_t>(outcap + 1);
        f.cursor = dcnt64 + 3;
        f.cap = outcap;
        f.kept = dk;
        CU(cudaEventRecord(ev[2], c->stream));
        k_cull_fused<<<(uint32_t)std::min<unsigned long long>(ntl, (unsigned long long)c->sm_count * 16), 256, 0, c->stream>>>(f, (uint32_t)ntl);
        CU(cudaEventRecord(ev[3], c->stream));
        c->be->launches += 2;
        CU(cudaGetLastError());
        c->be->d2h(hk.data(), dk, (size_t)nloc * 8);
        unsigned long long cur = 0;
        c->be->d2h(&cur, dcnt64 + 3, 8);
        stored = std::min(cur, outcap);
    } else {
        CU(cudaEventRecord(ev[1], c->stream));
        CU(cudaEventRecord(ev[2], c->stream));
        CU(cudaEventRecord(ev[3], c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    pcv_query_stats& q = c->qstats;
    cudaEventElapsedTime(&q.ms_device, ev[0], ev[3]);
    cudaEventElapsedTime(&q.ms_select, ev[0], ev[1]);
    cudaEventElapsedTime(&q.ms_cull, ev[2], ev[3]);
    for (auto& e : ev) cudaEventDestroy(e);
    q.kernel_launches = (uint32_t)(c->be->launches - l0);
    q.visited_pairs = npairs;
    for (uint32_t i = 0; i < nloc; ++i) {
        q.tested_points += ht[i];
        q.returned_points += hk[i];
        if (counts_out) counts_out[i] = hk[i];
        if (tested_out) tested_out[i] = ht[i];
    }
    q.stored_points = stored;
    q.algorithmic_bytes = h64[1] + 27ull * q.returned_points;
    return PCV_OK;
    API_CATCH
}

int pcv_lod_order(uint64_t seed, uint64_t id_high, uint64_t id_low, uint64_t n, uint64_t* new_order_out) {
    if (n && !new_order_out) return fail(PCV_ERR_INVALID, "null argument");
    if (n >= 0xFFFFFFFFull) return fail(PCV_ERR_UNSUPPORTED, "node too large");
    const uint64_t key = lod_node_key(seed, id_high, id_low);
    for (uint64_t i = 0; i < n; ++i) new_order_out[i] = lod_order(key, (uint32_t)n, (uint32_t)i);
    return PCV_OK;
}

int pcv_octree_shuffle_nodes(pcv_octree* o, uint64_t seed) {
    if (!o) return fail(PCV_ERR_INVALID, "null octree");
    API_TRY
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ensure_tables(o);
    if (o->n == 0) return PCV_OK;
    Scratch s(c);
    std::vector<QTile> tiles;
    std::vector<uint64_t> keys(o->nodes.size());
    for (size_t i = 0; i < o->nodes.size(); ++i) {
        keys[i] = lod_node_key(seed, o->nodes[i].id_high, o->nodes[i].id_low);
        if (o->nodes[i].num_points > 0) make_tiles(o, 0, (uint32_t)i, tiles);
    }
    LodArgs a{};
    a.nodes = (const QNode*)o->d_qnodes;
    a.tiles = s.upload(tiles.data(), tiles.size());
    a.keys = s.upload(keys.data(), keys.size());
    a.xyz = o->d_xyz;
    a.rgb = o->d_rgb;
    a.intensity = o->d_intensity;
    a.src = o->d_src;
    a.out_xyz = (uint8_t*)c->be->dmalloc(o->xyz_bytes + 32);
    a.out_rgb = (uint8_t*)c->be->dmalloc(o->n * 3);
    a.out_src = (uint32_t*)c->be->dmalloc(o->n * 4 + 64);
    a.out_intensity = o->d_intensity ? (float*)c->be->dmalloc(o->n * 4) : nullptr;
    k_lod_shuffle<<<(uint32_t)std::min<size_t>(tiles.size(), (size_t)c->sm_count * 16), 256, 0, c->stream>>>(a, (uint32_t)tiles.size());
    c->be->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->be->dfree(o->d_xyz);
    c->be->dfree(o->d_rgb);
    c->be->dfree(o->d_src);
    c->be->dfree(o->d_intensity);
    o->d_xyz = a.out_xyz;
    o->d_rgb = a.out_rgb;
    o->d_src = a.out_src;
    o->d_intensity = a.out_intensity;
    return PCV_OK;
    API_CATCH
}

int pcv_last_query_stats(pcv_ctx* c, pcv_query_stats* out) {
    if (!c || !out) return fail(PCV_ERR_INVALID, "null argument");
    *out = c->qstats;
    return PCV_OK;
}
int pcv_last_xray_stats(pcv_ctx* c, pcv_xray_stats* out) {
    if (!c || !out) return fail(PCV_ERR_INVALID, "null argument");
    *out = c->xstats;
    return PCV_OK;
}

// Shared by the X-ray entry points: the tile's location (Aabb, or Obb when a query_from_global transform is given), the nodes
// it intersects, their point tiles, and the kernel arguments that do not depend on the colouring strategy.
static void xray_prepare(pcv_octree* o, pcv_ctx* c, const double tmin[3], const double tmax[3], uint32_t w, uint32_t h, const double* qfg, Scratch& s,
                         XrayArgs& a, std::vector<uint32_t>* hit_nodes, size_t* ntiles_out = nullptr) {
    // location: Aabb(bbox), or Obb::from(bbox).transformed(query_from_global.inverse())  (xray generation.rs:471-477)
    pcv_location loc{};
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; ++a) {
        bmin[a] = std::fmin(tmin[a], tmax[a]);
        bmax[a] = std::fmax(tmin[a], tmax[a]);
    }
    if (qfg) {
        loc.kind = PCV_LOC_OBB;
        double ginv[7];  // global_from_query = query_from_global.inverse(): conjugate, t' = rot_inv * (-t)
        ginv[3] = -qfg[3];
        ginv[4] = -qfg[4];
        ginv[5] = -qfg[5];
        ginv[6] = qfg[6];
        const V3 nt = quat_rot(ginv, V3{-qfg[0], -qfg[1], -qfg[2]});
        ginv[0] = nt.x;
        ginv[1] = nt.y;
        ginv[2] = nt.z;
        // Obb::from(&aabb): centre = (min+max)*0.5, half = diag*0.5 (obb.rs:19-26); composed with the identity rotation
        const V3 centre{(bmin[0] + bmax[0]) * 0.5, (bmin[1] + bmax[1]) * 0.5, (bmin[2] + bmax[2]) * 0.5};
        const V3 sh = quat_rot(ginv, centre);
        double* q = loc.query_from_obb;
        q[0] = ginv[0] + sh.x;
        q[1] = ginv[1] + sh.y;
        q[2] = ginv[2] + sh.z;
        q[3] = ginv[3];
        q[4] = ginv[4];
        q[5] = ginv[5];
        q[6] = ginv[6];
        double* qi = loc.obb_from_query;
        qi[3] = -q[3];
        qi[4] = -q[4];
        qi[5] = -q[5];
        qi[6] = q[6];
        const V3 it = quat_rot(qi, V3{-q[0], -q[1], -q[2]});
        qi[0] = it.x;
        qi[1] = it.y;
        qi[2] = it.z;
        for (int a = 0; a < 3; ++a) loc.half_extent[a] = (bmax[a] - bmin[a]) * 0.5;
    } else {
        loc.kind = PCV_LOC_AABB;
        for (int a = 0; a < 3; ++a) {
            loc.aabb_min[a] = bmin[a];
            loc.aabb_max[a] = bmax[a];
        }
    }
    std::vector<QueryGeom> geoms{make_query_geom(loc)};
    std::vector<uint8_t> pass = run_sat(o, geoms, s, nullptr);
    std::vector<QTile> tiles;
    for (size_t i = 0; i < pass.size(); ++i)
        if (pass[i] && o->nodes[i].num_points > 0) {
            if (hit_nodes) hit_nodes->push_back((uint32_t)i);
            make_tiles(o, 0, (uint32_t)i, tiles);
        }
    a = XrayArgs{};
    a.geom = geoms[0];
    a.nodes = (const QNode*)o->d_qnodes;
    a.tiles = s.upload(tiles.data(), tiles.size());
    a.xyz = o->d_xyz;
    for (int k = 0; k < 3; ++k) {
        a.tmin[k] = bmin[k];
        a.tdiag[k] = bmax[k] - bmin[k];
        a.rdiag[k] = 1.0 / a.tdiag[k];
    }
    a.div_ok = div_known_ok(a.tdiag[0]) && div_known_ok(a.tdiag[1]) && div_known_ok(a.tdiag[2]) ? 1 : 0;
    a.has_q = qfg ? 1 : 0;
    if (qfg) memcpy(a.query_from_global, qfg, sizeof(double) * 7);
    a.w = w;
    a.h = h;
    if (ntiles_out) *ntiles_out = tiles.size();
    (void)c;
}

// The leaf tile with the XRay strategy.  The caller holds the context's lock and has selected its device.  The image goes to
// `rgba_out` (host, optional) and / or stays in `d_rgba_ext` (device, w * h * 4 bytes, optional: the quadtree driver keeps
// the tiles of a level resident for the parents).
static int xray_tile_core(pcv_octree* o, const double tmin[3], const double tmax[3], uint32_t w, uint32_t h, const double* qfg, uint8_t* rgba_out,
                          uint32_t* zbits_out, int* any_out, uint8_t* d_rgba_ext) {
    pcv_ctx* c = o->ctx;
    ensure_tables(o);
    c->xstats = pcv_xray_stats{};
    const uint64_t l0 = c->be->launches;
    Scratch s(c);
    XrayBinArgs bin{};
    XrayArgs& a = bin.x;
    std::vector<uint32_t> hit;  // nodes of the tile's location that hold points
    size_t ntiles = 0;
    xray_prepare(o, c, tmin, tmax, w, h, qfg, s, a, &hit, &ntiles);
    const size_t npix = (size_t)w * h;
    uint64_t pts = 0, bytes = 0;
    for (uint32_t k : hit) {
        const pcv_node_meta& m = o->nodes[k];
        pts += (uint64_t)m.num_points;
        bytes += (uint64_t)m.num_points * (3ull * (uint64_t)enc_bytes(m.position_encoding) + 3ull);
    }
    if (pts >= 0xFFFFFFFFull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-1 points in one X-ray tile");
    const uint32_t sw = (w + kXraySub - 1) / kXraySub, sh = (h + kXraySub - 1) / kXraySub, nsub = sw * sh;
    uint8_t grey[1026];
    grey[0] = 0;
    const double max_sat = std::log(1024.0);  // generation.rs:165-171
    for (int n = 1; n <= 1025; ++n) {
        const double v = (1. - std::log((double)n) / max_sat) * 255.;
        grey[n] = v != v || v <= 0.0 ? 0 : (v >= 255.0 ? 255 : (uint8_t)v);  // `as u8`
    }
    a.any = s.alloc<int>(1);
    bin.ntiles = (uint32_t)ntiles;
    bin.sub_w = sw;
    bin.sub_count = s.alloc<uint32_t>((size_t)nsub + 1);
    bin.sub_cursor = s.alloc<uint32_t>((size_t)nsub + 1);
    bin.keys = s.alloc<uint32_t>(std::max<uint64_t>(pts, 1));
    XraySubArgs sb{};
    sb.grey = s.upload(grey, 1026);
    sb.rgba = d_rgba_ext ? d_rgba_ext : s.alloc<uint8_t>(npix * 4);
    sb.zbits_out = zbits_out ? s.alloc<uint32_t>(npix * 32) : nullptr;
    sb.sub_w = sw;
    sb.w = w;
    sb.h = h;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, c->stream));
    k_fill_u32<<<(uint32_t)std::min<size_t>((npix + 255) / 256, (size_t)c->sm_count * 8), 256, 0, c->stream>>>((uint32_t*)sb.rgba, kXrayTransparent, npix);
    c->be->launches++;
    CU(cudaMemsetAsync(bin.sub_count, 0, ((size_t)nsub + 1) * 4, c->stream));
    CU(cudaMemsetAsync(bin.sub_cursor, 0, ((size_t)nsub + 1) * 4, c->stream));
    if (sb.zbits_out) CU(cudaMemsetAsync(sb.zbits_out, 0, npix * 128, c->stream));
    int any = 0;
    if (ntiles) {
        const uint32_t grid = (uint32_t)std::min<size_t>(ntiles, (size_t)c->sm_count * 16);
        k_xray_bin<0><<<grid, 256, 0, c->stream>>>(bin);
        unsigned long long* dtot = s.alloc<unsigned long long>(1);
        k_scan_u32<<<1, 1024, 0, c->stream>>>(bin.sub_count, nsub + 1, dtot);  // exclusive offsets; entry nsub = total
        k_xray_bin<1><<<grid, 256, 0, c->stream>>>(bin);
        c->be->launches += 3;
        CU(cudaGetLastError());
        // one block per sub-tile, enqueued without looking at the counts (an empty sub-tile's block leaves at once): no host
        // round trip inside a tile
        sb.sub_id = nullptr;
        sb.sub_off = bin.sub_count;
        sb.keys = bin.keys;
        const size_t sm = (size_t)kXraySub * kXraySub * 128;
        k_xray_subtile<<<nsub, 512, sm, c->stream>>>(sb);
        c->be->launches++;
        CU(cudaEventRecord(e1, c->stream));
        CU(cudaGetLastError());
        unsigned long long total = 0;
        c->be->d2h(&total, dtot, 8);
        any = total != 0;
    } else {
        CU(cudaEventRecord(e1, c->stream));
    }
    CU(cudaGetLastError());
    if (rgba_out)
        c->be->d2h(rgba_out, sb.rgba, npix * 4);
    else
        CU(cudaStreamSynchronize(c->stream));  // the events below and the scratch buffers
    if (zbits_out) c->be->d2h(zbits_out, sb.zbits_out, npix * 128);
    if (any_out) *any_out = any;
    cudaEventElapsedTime(&c->xstats.ms_device, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->xstats.kernel_launches = (uint32_t)(c->be->launches - l0);
    c->xstats.points = pts;
    c->xstats.algorithmic_bytes = bytes + 4ull * npix;
    return PCV_OK;
}

int pcv_xray_tile(const pcv_octree* oc, const double tmin[3], const double tmax[3], uint32_t w, uint32_t h, const double* qfg, uint8_t* rgba_out,
                  uint32_t* zbits_out, int* any_out) {
    if (!oc || !tmin || !tmax || !rgba_out || w == 0 || h == 0) return fail(PCV_ERR_INVALID, "null argument or empty image");
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    return xray_tile_core(o, tmin, tmax, w, h, qfg, rgba_out, zbits_out, any_out, nullptr);
    API_CATCH
}

}  // extern "C"

// The other ColoringStrategyKinds of xray_from_points (xray/src/generation.rs:76-97, 200-405).  bin_size == 0: Binning = None;
// bin_size > 0 (strategies Colored and ColoredWithIntensity): Binning = Some(("intensity", bin_size)), the columns are
// aggregated per (pixel, bin) in two device hash tables (xray_pyramid.h) and the pixel is the mean of its bins' means.
// Same contract as xray_tile_core for the lock and the two outputs.
static int xray_tile_attr_core(pcv_octree* o, const double tmin[3], const double tmax[3], uint32_t w, uint32_t h, const double* qfg, int mode, float p0,
                               float p1, int colormap, double bin_size, uint8_t* rgba_out, int* any_out, uint8_t* d_rgba_ext) {
    pcv_ctx* c = o->ctx;
    ensure_tables(o);
    Scratch s(c);
    const bool binned = bin_size != 0.0;
    XrayBinnedArgs bb{};
    XrayAttrArgs b{};
    XrayArgs& xa = binned ? bb.x : b.x;
    size_t ntiles = 0;
    std::vector<uint32_t> hit;
    xray_prepare(o, c, tmin, tmax, w, h, qfg, s, xa, &hit, &ntiles);
    const size_t npix = (size_t)w * h;
    xa.any = s.alloc<int>(1);
    b.rgb = o->d_rgb;
    b.intensity = o->d_intensity;
    b.count = s.alloc<uint32_t>(npix);
    b.z0 = (std::fmin(tmin[2], tmax[2]) + std::fmax(tmin[2], tmax[2])) * 0.5;
    CU(cudaMemsetAsync(xa.any, 0, 4, c->stream));
    CU(cudaMemsetAsync(b.count, 0, npix * 4, c->stream));
    if (mode == PCV_XRAY_HEIGHT_STDDEV) {
        b.dsum = s.alloc<double>(npix * 2);
        CU(cudaMemsetAsync(b.dsum, 0, npix * 16, c->stream));
    } else {
        const size_t per = mode == PCV_XRAY_COLORED ? 4 : 1;
        b.sum = s.alloc<float>(npix * per);
        CU(cudaMemsetAsync(b.sum, 0, npix * per * 4, c->stream));
    }
    uint8_t* drgba = d_rgba_ext ? d_rgba_ext : s.alloc<uint8_t>(npix * 4);
    int* derr = nullptr;
    if (ntiles && !binned) {
        if (mode == PCV_XRAY_COLORED)
            k_xray_accum_attr<1><<<(uint32_t)ntiles, 256, 0, c->stream>>>(b);
        else if (mode == PCV_XRAY_INTENSITY)
            k_xray_accum_attr<2><<<(uint32_t)ntiles, 256, 0, c->stream>>>(b);
        else
            k_xray_accum_attr<3><<<(uint32_t)ntiles, 256, 0, c->stream>>>(b);
        c->be->launches++;
    } else if (ntiles) {
        uint64_t pts = 0;
        for (uint32_t k : hit) pts += (uint64_t)o->nodes[k].num_points;
        BinnedTables& t = bb.t;
        t.bin_cap = 1u << 20;
        t.col_cap = 2 * pts + 1024;
        t.ncomp = mode == PCV_XRAY_COLORED ? 3 : 1;
        t.bin_keys = s.alloc<uint64_t>(t.bin_cap);
        t.col_keys = s.alloc<uint64_t>(t.col_cap);
        t.col_sum = s.alloc<float>(t.col_cap * (size_t)t.ncomp);
        t.col_count = s.alloc<uint32_t>(t.col_cap);
        t.err = derr = s.alloc<int>(1);
        const uint32_t fgrid = (uint32_t)c->sm_count * 8;
        k_fill_u64<<<fgrid, 256, 0, c->stream>>>(t.bin_keys, kBinEmpty, t.bin_cap);
        k_fill_u64<<<fgrid, 256, 0, c->stream>>>(t.col_keys, kColEmpty, t.col_cap);
        CU(cudaMemsetAsync(t.col_sum, 0, t.col_cap * (size_t)t.ncomp * 4, c->stream));
        CU(cudaMemsetAsync(t.col_count, 0, t.col_cap * 4, c->stream));
        CU(cudaMemsetAsync(t.err, 0, 4, c->stream));
        bb.rgb = o->d_rgb;
        bb.intensity = o->d_intensity;
        bb.bin_size = bin_size;
        if (mode == PCV_XRAY_COLORED)
            k_xray_binned_insert<1><<<(uint32_t)ntiles, 256, 0, c->stream>>>(bb);
        else
            k_xray_binned_insert<2><<<(uint32_t)ntiles, 256, 0, c->stream>>>(bb);
        // per (pixel, bin) mean -> the pixel's sum over bins and its number of bins, in the layout k_xray_resolve_attr reads
        k_xray_binned_reduce<<<fgrid, 256, 0, c->stream>>>(t, b.sum, mode == PCV_XRAY_COLORED ? 4 : 1, b.count);
        c->be->launches += 4;
    }
    k_xray_resolve_attr<<<(uint32_t)((npix + 255) / 256), 256, 0, c->stream>>>(mode, p0, p1, colormap, b.sum, b.dsum, b.count, (uint32_t)npix, drgba);
    c->be->launches++;
    CU(cudaGetLastError());
    int any = 0;
    c->be->d2h(&any, xa.any, 4);
    if (derr) {
        int err = 0;
        c->be->d2h(&err, derr, 4);
        if (err) return fail(PCV_ERR_UNSUPPORTED, err == 1 ? "more than 2^20 distinct bins in one X-ray tile" : "X-ray column table full");
    }
    if (rgba_out) c->be->d2h(rgba_out, drgba, npix * 4);
    if (any_out) *any_out = any;
    return PCV_OK;
}

static int xray_attr_check(const pcv_octree* oc, int mode, double bin_size) {
    if (mode < PCV_XRAY_COLORED || mode > PCV_XRAY_HEIGHT_STDDEV) return fail(PCV_ERR_INVALID, "unknown colouring strategy %d", mode);
    if (mode == PCV_XRAY_INTENSITY && !oc->d_intensity)
        return fail(PCV_ERR_INVALID, "Coloring by intensity was requested, but point data without intensity found.");
    if (bin_size != 0.0) {
        if (mode == PCV_XRAY_HEIGHT_STDDEV) return fail(PCV_ERR_INVALID, "the height-stddev strategy has no binning (xray/src/generation.rs:76-84)");
        if (!(bin_size == bin_size)) return fail(PCV_ERR_INVALID, "bin size is NaN");
        if (!oc->d_intensity) return fail(PCV_ERR_INVALID, "Binning attribute needs to be available in points batch.");
    }
    return PCV_OK;
}

int pcv_xray_tile_attr(const pcv_octree* oc, const double tmin[3], const double tmax[3], uint32_t w, uint32_t h, const double* qfg, int mode, float p0,
                       float p1, int colormap, uint8_t* rgba_out, int* any_out) {
    if (!oc || !tmin || !tmax || !rgba_out || w == 0 || h == 0) return fail(PCV_ERR_INVALID, "null argument or empty image");
    if (int rc = xray_attr_check(oc, mode, 0.0)) return rc;
    API_TRY
    pcv_octree* o = const_cast<pcv_octree*>(oc);
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    return xray_tile_attr_core(o, tmin, tmax, w, h, qfg, mode, p0, p1, colormap, 0.0, rgba_out, any_out, nullptr);
    API_CATCH
}

// ---- /nodes_data reply blob (octree_web_viewer/src/backend.rs:66-75 pad, :92-165 get_nodes_data) ----------------
// Per requested node, in request order: cube min x, y, z (f64 LE), edge length (f64), num_points (u32), bytes per
// coordinate (u8), zeros up to a multiple of 8, the node's position bytes, padding, the node's colour bytes, padding.
// Octree::get_node_data fails with NodeNotFound for ids without files - unknown ids and nodes with zero points, whose
// files were deleted (data_provider/on_disk.rs:51-68, node_writer.rs:78-89): PCV_ERR_NOT_FOUND here.
// One gather kernel + one device-to-host copy serve the whole request from HBM.
int pcv_nodes_data_blob(const pcv_octree* o, const uint64_t* ids_hi_lo, uint32_t num_nodes, void* out, uint64_t cap, uint64_t* size_out) {
    if (!o || (num_nodes && !ids_hi_lo) || !size_out) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    auto pad8 = [](uint64_t v) { return (v + 7) & ~(uint64_t)7; };
    std::vector<BlobItem> items;
    std::vector<uint64_t> header_at(num_nodes);
    std::vector<int> which(num_nodes);
    uint64_t size = 0;
    for (uint32_t k = 0; k < num_nodes; ++k) {
        const int i = o->find(ids_hi_lo[2 * k], ids_hi_lo[2 * k + 1]);
        if (i < 0 || o->nodes[i].num_points == 0)
            return fail(PCV_ERR_NOT_FOUND, "Could not get node %s.", node_name(ids_hi_lo[2 * k], ids_hi_lo[2 * k + 1]).c_str());
        const pcv_node_meta& m = o->nodes[i];
        which[k] = i;
        header_at[k] = size;
        size += pad8(8 * 4 + 4 + 1);
        const uint64_t n = (uint64_t)m.num_points, pb = n * 3 * (uint64_t)enc_bytes(m.position_encoding), cb = n * 3;
        for (uint64_t o2 = 0; o2 < pb; o2 += kBlobSeg) items.push_back(BlobItem{m.xyz_byte_offset + o2, size + o2, (uint32_t)std::min<uint64_t>(kBlobSeg, pb - o2), 0});
        size += pad8(pb);
        for (uint64_t o2 = 0; o2 < cb; o2 += kBlobSeg) items.push_back(BlobItem{3 * m.point_offset + o2, size + o2, (uint32_t)std::min<uint64_t>(kBlobSeg, cb - o2), 1});
        size += pad8(cb);
    }
    *size_out = size;
    if (!out) return PCV_OK;  // size query
    if (cap < size) return fail(PCV_ERR_INVALID, "reply buffer too small: %llu < %llu", (unsigned long long)cap, (unsigned long long)size);
    if (size == 0) return PCV_OK;
    pcv_ctx* c = o->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CudaBackend& be = *c->be;
    uint8_t* d_blob = (uint8_t*)be.dmalloc(size);
    BlobItem* d_items = (BlobItem*)be.dmalloc(items.size() * sizeof(BlobItem));
    try {
        CU(cudaMemsetAsync(d_blob, 0, size, c->stream));  // padding bytes are zeros
        be.h2d(d_items, items.data(), items.size() * sizeof(BlobItem));
        k_blob_gather<<<(uint32_t)items.size(), 256, 0, c->stream>>>(d_items, o->d_xyz, o->d_rgb, d_blob);
        ++be.launches;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, d_blob, size, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } catch (...) {
        be.dfree(d_blob);
        be.dfree(d_items);
        throw;
    }
    be.dfree(d_blob);
    be.dfree(d_items);
    // the 40-byte headers are host data (node table): written after the copy
    uint8_t* ob = (uint8_t*)out;
    for (uint32_t k = 0; k < num_nodes; ++k) {
        const pcv_node_meta& m = o->nodes[which[k]];
        uint8_t* h = ob + header_at[k];
        std::memcpy(h, m.cube_min, 24);
        std::memcpy(h + 24, &m.cube_edge, 8);
        const uint32_t n32 = (uint32_t)m.num_points;  // `as u32`
        std::memcpy(h + 32, &n32, 4);
        h[36] = (uint8_t)enc_bytes(m.position_encoding);
        h[37] = h[38] = h[39] = 0;
    }
    return PCV_OK;
    API_CATCH
}
