// shard_api.inl — multi-GPU sharding entry points (included at the end of pcv_api.cu).

namespace {

PrefixArgs make_prefix_args(const pcv_points* dp, double resolution, const double bmin_in[3], const double bmax_in[3], uint32_t k) {
    if (k < 1 || k > 3) throw BuildError(PCV_ERR_INVALID, "prefix levels k must be 1..3");
    if (!(resolution > 0.0)) throw BuildError(PCV_ERR_INVALID, "resolution must be > 0");
    PrefixArgs a{};
    a.pts = view_of(dp);
    double bmin[3], bmax[3];
    for (int i = 0; i < 3; ++i) {
        bmin[i] = std::fmin(bmin_in[i], bmax_in[i]);
        bmax[i] = std::fmax(bmin_in[i], bmax_in[i]);
        a.root_min[i] = bmin[i];
    }
    const double E = std::fmax(std::fmax(bmax[0] - bmin[0], bmax[1] - bmin[1]), bmax[2] - bmin[2]);
    a.lv = make_level_table(E, resolution);
    a.k = (int)k;
    a.nbins = 1 << (3 * k);
    return a;
}

}  // namespace

extern "C" {

static int prefix_histogram_impl(pcv_ctx* c, const pcv_points* dp, double resolution, const double bmin[3], const double bmax[3], uint32_t k,
                                 uint64_t* counts_out, double* data_min, double* data_max) {
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    PrefixArgs a = make_prefix_args(dp, resolution, bmin, bmax, k);
    std::vector<unsigned long long> h((size_t)a.nbins, 0);
    if (data_min)
        for (int i = 0; i < 3; ++i) data_min[i] = data_max[i] = 0.0;  // Aabb::zero for no points (generation.rs:269)
    if (a.pts.n) {
        unsigned long long* d = (unsigned long long*)c->be->dmalloc((size_t)a.nbins * 8);
        CU(cudaMemsetAsync(d, 0, (size_t)a.nbins * 8, c->stream));
        const int blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * 8, (a.pts.n + 255) / 256);
        double* d_part = data_min ? (double*)c->be->dmalloc((size_t)blocks * 6 * 8) : nullptr;
        // keep the per-point cells for the pack that follows (freed by the next histogram call or with the context)
        c->be->dfree(c->shard_cells);
        c->shard_cells = (uint16_t*)c->be->dmalloc(a.pts.n * 2);
        c->shard_cells_x = a.pts.x;
        c->shard_cells_n = a.pts.n;
        c->shard_cells_k = k;
        c->shard_cells_geom[0] = resolution;
        for (int i = 0; i < 3; ++i) c->shard_cells_geom[1 + i] = bmin[i], c->shard_cells_geom[4 + i] = bmax[i];
        k_prefix_hist<<<blocks, 256, (size_t)a.nbins * 4, c->stream>>>(a, d, c->shard_cells, d_part);
        c->be->launches++;
        CU(cudaGetLastError());
        c->be->d2h(h.data(), d, (size_t)a.nbins * 8);
        c->be->dfree(d);
        if (d_part) {
            std::vector<double> part((size_t)blocks * 6);
            c->be->d2h(part.data(), d_part, part.size() * 8);
            c->be->dfree(d_part);
            for (int i = 0; i < 3; ++i) data_min[i] = part[i], data_max[i] = part[3 + i];
            for (int b = 1; b < blocks; ++b)
                for (int i = 0; i < 3; ++i) {
                    data_min[i] = std::fmin(data_min[i], part[(size_t)b * 6 + i]);
                    data_max[i] = std::fmax(data_max[i], part[(size_t)b * 6 + 3 + i]);
                }
        }
    }
    for (int i = 0; i < a.nbins; ++i) counts_out[i] = h[(size_t)i];
    return PCV_OK;
}

int pcv_prefix_histogram_device(pcv_ctx* c, const pcv_points* dp, double resolution, const double bmin[3], const double bmax[3], uint32_t k,
                                uint64_t* counts_out) {
    if (!c || !dp || !bmin || !bmax || !counts_out) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    return prefix_histogram_impl(c, dp, resolution, bmin, bmax, k, counts_out, nullptr, nullptr);
    API_CATCH
}

int pcv_prefix_histogram_bbox_device(pcv_ctx* c, const pcv_points* dp, double resolution, const double bmin[3], const double bmax[3], uint32_t k,
                                     uint64_t* counts_out, double data_min[3], double data_max[3]) {
    if (!c || !dp || !bmin || !bmax || !counts_out || !data_min || !data_max) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    return prefix_histogram_impl(c, dp, resolution, bmin, bmax, k, counts_out, data_min, data_max);
    API_CATCH
}

static void attach_cells(pcv_ctx* c, PackArgs& a, double resolution, const double bmin[3], const double bmax[3], uint32_t k) {
    a.cells = nullptr;
    a.cell_shift = 0;
    if (!c->shard_cells || c->shard_cells_x != a.p.pts.x || c->shard_cells_n != a.p.pts.n || c->shard_cells_k < k) return;
    if (c->shard_cells_geom[0] != resolution) return;
    for (int i = 0; i < 3; ++i)
        if (c->shard_cells_geom[1 + i] != bmin[i] || c->shard_cells_geom[4 + i] != bmax[i]) return;
    a.cells = c->shard_cells;
    a.cell_shift = 3 * (c->shard_cells_k - k);
}
// The cells of a histogram call serve exactly ONE pack: a caller that refills the same buffer (a caching allocator hands out the same
// address and size every step) must not find the previous step's cells.  Called after the pack's kernels have been enqueued.
static void release_cells(pcv_ctx* c) {
    if (!c->shard_cells) return;
    c->be->dfree(c->shard_cells);  // stream-ordered: the pack that reads them was enqueued before
    c->shard_cells = nullptr;
    c->shard_cells_x = nullptr;
    c->shard_cells_n = 0;
    c->shard_cells_k = 0;
}

int pcv_prefix_pack_device(pcv_ctx* c, const pcv_points* dp, const uint64_t* dgidx, uint64_t gidx_base, double resolution, const double bmin[3],
                           const double bmax[3], uint32_t k, const int32_t* cell_to_rank, uint32_t nranks, double* out_xyz, uint8_t* out_rgb,
                           float* out_intensity, uint64_t* out_idx, uint64_t* rank_counts_out) {
    if (!c || !dp || !bmin || !bmax || !cell_to_rank || !out_xyz || !out_rgb || !out_idx || !rank_counts_out)
        return fail(PCV_ERR_INVALID, "null argument");
    if (nranks == 0 || nranks > (uint32_t)kMaxRanks) return fail(PCV_ERR_INVALID, "nranks must be 1..%d", kMaxRanks);
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    PackArgs a{};
    a.p = make_prefix_args(dp, resolution, bmin, bmax, k);
    for (uint32_t r = 0; r < nranks; ++r) rank_counts_out[r] = 0;
    const uint64_t n = a.p.pts.n;
    if (n == 0) return PCV_OK;
    if (n >= 0xFFFFFFFFull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-2 points per context");
    if (!a.p.pts.rgb) return fail(PCV_ERR_INVALID, "color is mandatory");
    for (int i = 0; i < a.p.nbins; ++i)
        if (cell_to_rank[i] < 0 || cell_to_rank[i] >= (int32_t)nranks) return fail(PCV_ERR_INVALID, "cell_to_rank[%d] out of range", i);
    Scratch s(c);
    a.cell_to_rank = s.upload(cell_to_rank, (size_t)a.p.nbins);
    a.nranks = nranks;
    a.ntiles = (uint32_t)((n + kPackTile - 1) / kPackTile);
    a.dest = s.alloc<uint8_t>(n);
    a.counts = s.alloc<uint32_t>((size_t)nranks * a.ntiles);
    a.gidx_in = dgidx;
    a.gidx_base = gidx_base;
    a.out_xyz = out_xyz;
    a.out_rgb = out_rgb;
    a.out_intensity = a.p.pts.intensity ? out_intensity : nullptr;
    a.out_idx = out_idx;
    attach_cells(c, a, resolution, bmin, bmax, k);
    k_pack_count<<<a.ntiles, 256, 0, c->stream>>>(a);
    unsigned long long* dtot = s.alloc<unsigned long long>(1);
    // per-rank totals before the scan overwrites the counts: read back the flat array's rank boundaries afterwards
    k_scan_u32<<<1, 1024, 0, c->stream>>>(a.counts, nranks * a.ntiles, dtot);
    k_pack_scatter<<<a.ntiles, 256, 0, c->stream>>>(a);
    c->be->launches += 3;
    CU(cudaGetLastError());
    release_cells(c);
    // rank r's first slot = exclusive prefix at (r, tile 0); counts follow from consecutive starts
    std::vector<uint32_t> starts(nranks);
    for (uint32_t r = 0; r < nranks; ++r) c->be->d2h(&starts[r], a.counts + (size_t)r * a.ntiles, 4);
    for (uint32_t r = 0; r < nranks; ++r) rank_counts_out[r] = (r + 1 < nranks ? starts[r + 1] : (uint32_t)n) - starts[r];
    return PCV_OK;
    API_CATCH
}

// ---- peer memory for the fused pack + exchange (CUDA IPC; one process per GPU) ----------------------------------------
int pcv_ipc_alloc(pcv_ctx* c, uint64_t bytes, void** dev_ptr, uint8_t handle_out[64]) {
    if (!c || !dev_ptr || !handle_out) return fail(PCV_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes ? bytes : 256));  // plain cudaMalloc: stream-ordered pool memory cannot be exported
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        throw BuildError(PCV_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, 64);
    *dev_ptr = p;
    return PCV_OK;
    API_CATCH
}
int pcv_ipc_free(pcv_ctx* c, void* dev_ptr) {
    if (!c) return fail(PCV_ERR_INVALID, "null context");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (dev_ptr) CU(cudaFree(dev_ptr));
    return PCV_OK;
    API_CATCH
}
int pcv_ipc_open(pcv_ctx* c, const uint8_t handle[64], void** dev_ptr) {
    if (!c || !handle || !dev_ptr) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PCV_OK;
    API_CATCH
}
int pcv_ipc_close(pcv_ctx* c, void* dev_ptr) {
    if (!c) return fail(PCV_ERR_INVALID, "null context");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (dev_ptr) CU(cudaIpcCloseMemHandle(dev_ptr));
    return PCV_OK;
    API_CATCH
}

int pcv_prefix_pack_exchange_device(pcv_ctx* c, const pcv_points* dp, const uint64_t* dgidx, uint64_t gidx_base, double resolution,
                                    const double bmin[3], const double bmax[3], uint32_t k, const int32_t* cell_to_rank, uint32_t nranks,
                                    const uint64_t* dst_first, void* const* dst_x, void* const* dst_y, void* const* dst_z, void* const* dst_index,
                                    void* const* dst_intensity, void* const* dst_colour, uint64_t* rank_counts_out) {
    if (!c || !dp || !bmin || !bmax || !cell_to_rank || !dst_first || !dst_x || !dst_y || !dst_z || !dst_index || !dst_colour || !rank_counts_out)
        return fail(PCV_ERR_INVALID, "null argument");
    if (nranks == 0 || nranks > (uint32_t)kMaxRanks) return fail(PCV_ERR_INVALID, "nranks must be 1..%d", kMaxRanks);
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    PackArgs a{};
    a.p = make_prefix_args(dp, resolution, bmin, bmax, k);
    for (uint32_t r = 0; r < nranks; ++r) rank_counts_out[r] = 0;
    const uint64_t n = a.p.pts.n;
    if (n == 0) return PCV_OK;
    if (n >= 0xFFFFFFFFull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-2 points per context");
    if (!a.p.pts.rgb) return fail(PCV_ERR_INVALID, "color is mandatory");
    if (a.p.pts.intensity && !dst_intensity) return fail(PCV_ERR_INVALID, "intensity destinations missing");
    for (int i = 0; i < a.p.nbins; ++i)
        if (cell_to_rank[i] < 0 || cell_to_rank[i] >= (int32_t)nranks) return fail(PCV_ERR_INVALID, "cell_to_rank[%d] out of range", i);
    Scratch s(c);
    a.cell_to_rank = s.upload(cell_to_rank, (size_t)a.p.nbins);
    a.nranks = nranks;
    a.ntiles = (uint32_t)((n + kPackTile - 1) / kPackTile);
    a.dest = s.alloc<uint8_t>(n);
    a.counts = s.alloc<uint32_t>((size_t)nranks * a.ntiles);
    a.gidx_in = dgidx;
    a.gidx_base = gidx_base;
    PeerTable pt{};
    for (uint32_t r = 0; r < nranks; ++r) {
        pt.x[r] = (double*)dst_x[r];
        pt.y[r] = (double*)dst_y[r];
        pt.z[r] = (double*)dst_z[r];
        pt.idx[r] = (uint64_t*)dst_index[r];
        pt.intensity[r] = dst_intensity ? (float*)dst_intensity[r] : nullptr;
        pt.col[r] = (uint32_t*)dst_colour[r];
        pt.first[r] = dst_first[r];
    }
    const PeerTable* d_pt = s.upload(&pt, 1);
    attach_cells(c, a, resolution, bmin, bmax, k);
    k_pack_count<<<a.ntiles, 256, 0, c->stream>>>(a);
    unsigned long long* dtot = s.alloc<unsigned long long>(1);
    k_scan_u32<<<1, 1024, 0, c->stream>>>(a.counts, nranks * a.ntiles, dtot);
    static const bool v1 = std::getenv("PCV_EXCHANGE_V1") != nullptr;  // diagnostic: the unsorted variant
    if (v1)
        k_pack_exchange<<<a.ntiles, 256, 0, c->stream>>>(a, d_pt);
    else
        k_pack_exchange_sorted<<<a.ntiles, 256, 0, c->stream>>>(a, d_pt);
    c->be->launches += 3;
    CU(cudaGetLastError());
    release_cells(c);
    std::vector<uint32_t> starts(nranks);
    for (uint32_t r = 0; r < nranks; ++r) c->be->d2h(&starts[r], a.counts + (size_t)r * a.ntiles, 4);  // synchronises: the stores are out
    for (uint32_t r = 0; r < nranks; ++r) rank_counts_out[r] = (r + 1 < nranks ? starts[r + 1] : (uint32_t)n) - starts[r];
    return PCV_OK;
    API_CATCH
}

int pcv_unpack_colours_device(pcv_ctx* c, const uint32_t* dev_colour, uint64_t n, uint8_t* dev_rgb) {
    if (!c || (n && (!dev_colour || !dev_rgb))) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    if (n == 0) return PCV_OK;
    const int blocks = (int)std::min<uint64_t>((uint64_t)c->sm_count * 16, (n + 255) / 256);
    k_unpack_colours<<<blocks, 256, 0, c->stream>>>(dev_colour, n, dev_rgb);
    c->be->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return PCV_OK;
    API_CATCH
}

int pcv_build_octree_sharded_device(pcv_ctx* c, const pcv_points* dp, double resolution, const double bmin[3], const double bmax[3], uint32_t k,
                                    const uint64_t* prefix_counts, pcv_octree** out) {
    if (!c || !dp || !bmin || !bmax || !prefix_counts || !out) return fail(PCV_ERR_INVALID, "null argument");
    if (k < 1 || k > 3) return fail(PCV_ERR_INVALID, "prefix levels k must be 1..3");
    *out = nullptr;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ShardSpec sp;
    sp.k = (int)k;
    sp.counts = prefix_counts;
    return build_impl(c, view_of(dp), resolution, bmin, bmax, out, &sp);
    API_CATCH
}

struct pcv_shard_send {
    pcv_ctx* ctx = nullptr;
    uint64_t n = 0;
    bool wide = false, has_intensity = false;
    int G0 = 2, nbins = 64;
    uint32_t ntiles = 0;
    void* rec = nullptr;
    uint32_t* col = nullptr;
    uint8_t* dig = nullptr;
    const float* intensity = nullptr;  // the caller's array (not owned)
    uint32_t* tile_counts = nullptr;
    uint8_t* dest = nullptr;
    uint32_t* tile_active = nullptr;
    LevelTable lv;
    double bmin[3] = {0, 0, 0}, root_edge = 0, resolution = 0;
    bool fused_done = false;     // pcv_shard_pass_device ran: `dig` (the level-2 cell of every local point) is what stays for provenance
    std::vector<uint64_t> bins;  // points per digit of the local points
    std::vector<void*> owned;
};

int pcv_shard_ingest_device(pcv_ctx* c, const pcv_points* dp, double resolution, const double bmin_in[3], const double bmax_in[3], uint32_t k,
                            uint64_t* counts_out, pcv_shard_send** out) {
    if (!c || !dp || !bmin_in || !bmax_in || !counts_out || !out) return fail(PCV_ERR_INVALID, "null argument");
    if (k < 1 || k > 2) return fail(PCV_ERR_INVALID, "prefix levels k must be 1 or 2");
    *out = nullptr;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; ++a) {
        bmin[a] = std::fmin(bmin_in[a], bmax_in[a]);
        bmax[a] = std::fmax(bmin_in[a], bmax_in[a]);
    }
    const PointsView v = view_of(dp);
    const uint64_t n = v.n;
    for (uint64_t i = 0; i < ((uint64_t)1 << (3 * k)); ++i) counts_out[i] = 0;
    if (n >= 0xFFFFFFFFull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-2 points per context");
    if (n && !v.rgb) return fail(PCV_ERR_INVALID, "color is mandatory");
    const double E = std::fmax(std::fmax(bmax[0] - bmin[0], bmax[1] - bmin[1]), bmax[2] - bmin[2]);
    const LevelTable lv = make_level_table(E, resolution, bmin);
    if ((int)k > lv.last_level) return fail(PCV_ERR_INVALID, "prefix levels exceed the depth of the octree");
    std::unique_ptr<pcv_shard_send> sd(new pcv_shard_send());
    sd->ctx = c;
    sd->n = n;
    sd->has_intensity = v.intensity != nullptr;
    sd->intensity = v.intensity;
    for (int L = 1; L <= lv.last_level; ++L) sd->wide = sd->wide || lv.enc[L] == ENC_F64;
    sd->G0 = std::min(2, lv.last_level);
    sd->nbins = 1 << (3 * sd->G0);
    sd->lv = lv;
    sd->root_edge = E;
    sd->resolution = resolution;
    for (int a = 0; a < 3; ++a) sd->bmin[a] = bmin[a];
    if (n == 0) {
        *out = sd.release();
        return PCV_OK;
    }
    CudaBackend& be = *c->be;
    auto dalloc = [&](size_t bytes) {
        void* p = be.dmalloc(bytes);
        sd->owned.push_back(p);
        return p;
    };
    const char* tv = std::getenv("PCV_TIMING");
    const bool fine = tv && tv[0] == '2';
    auto tp0 = std::chrono::steady_clock::now();
    double tms[5] = {0, 0, 0, 0, 0};
    auto tmark = [&](int i) {
        if (!fine) return;
        const auto t = std::chrono::steady_clock::now();
        tms[i] = std::chrono::duration<double, std::milli>(t - tp0).count();
        tp0 = t;
    };
    const size_t rec_bytes = sd->wide ? sizeof(RecW) : sizeof(RecN);
    const uint32_t nt = (uint32_t)((n + kTilePoints - 1) / kTilePoints), nch = (nt + kChunkTiles - 1) / kChunkTiles;
    sd->ntiles = nt;
    sd->rec = dalloc((size_t)n * rec_bytes + 64);
    sd->col = (uint32_t*)dalloc((size_t)n * 4 + 64);
    sd->dig = (uint8_t*)dalloc((size_t)n + 64);
    sd->dest = (uint8_t*)dalloc((size_t)n + 64);
    sd->tile_counts = (uint32_t*)dalloc((size_t)nt * 64 * 4);
    IngestArgs ia{};
    ia.pts = v;
    ia.rec_out = sd->rec;
    ia.col_out = sd->col;
    ia.dig_out = sd->dig;
    ia.G0 = sd->G0;
    ia.wide = sd->wide;
    ia.ntiles = nt;
    ia.lv = lv;
    for (int a = 0; a < 3; ++a) ia.root_min[a] = bmin[a];
    tmark(0);
    be.ingest(ia);
    if (fine) CU(cudaStreamSynchronize(c->stream));
    tmark(1);
    // per-tile digit histogram + exclusive prefix over the tiles (one "node": all local points), as a partition pass would
    Scratch s(c);
    BuildState hs{};
    hs.nnodes = 1;
    hs.pass[0].nactive = 1;
    hs.pass[0].ntiles = nt;
    hs.pass[0].nchunks = nch;
    hs.pass[0].npoints = n;
    BuildState* dst = s.upload(&hs, 1);
    ActiveDesc a0{};
    a0.count = n;
    a0.nchunks = nch;
    const ActiveDesc* dact = s.upload(&a0, 1);
    std::vector<ChunkDesc> c0;
    for (uint32_t o = 0; o < nt; o += kChunkTiles) c0.push_back(ChunkDesc{o, std::min(kChunkTiles, nt - o), 0u, o == 0 ? 1u : 0u});
    const ChunkDesc* dch = s.upload(c0.data(), c0.size());
    PassArgs pa{};
    pa.pass = 0;
    pa.G = sd->G0;
    pa.nbins = sd->nbins;
    pa.dig_in = sd->dig;
    pa.st = dst;
    pa.active = dact;
    pa.chunks = dch;
    pa.tile_counts = sd->tile_counts;
    sd->tile_active = (uint32_t*)dalloc((size_t)nt * 4);
    pa.tile_active = sd->tile_active;
    pa.chunk_sums = s.alloc<uint32_t>((size_t)nch * 64);
    pa.node_bins = s.alloc<uint64_t>(64);
    pa.cap_active = 1;
    pa.cap_chunks = nch;
    pa.cap_tiles = nt;
    tmark(2);
    be.hist_scan(pa);
    std::vector<uint64_t> bins(64, 0);
    be.d2h(bins.data(), pa.node_bins, (size_t)sd->nbins * 8);
    tmark(3);
    if (fine) fprintf(stderr, "[pcv ingest dev%d] allocs %.2f  ingest kernel %.2f  uploads %.2f  hist+scan+readback %.2f ms\n", c->device, tms[0], tms[1], tms[2], tms[3]);
    const int shift = 3 * (sd->G0 - (int)k);  // digits carry G0 levels; the shard level may be shallower
    if (shift < 0) return fail(PCV_ERR_INVALID, "prefix levels exceed the levels of the first pass");
    for (int d = 0; d < sd->nbins; ++d) counts_out[d >> shift] += bins[(size_t)d];
    sd->bins = bins;
    *out = sd.release();
    return PCV_OK;
    API_CATCH
}

void pcv_shard_send_free(pcv_shard_send* s) {
    if (!s) return;
    pcv_ctx* c = s->ctx;
    {
        std::lock_guard<std::mutex> g(c->mu);
        cudaSetDevice(c->device);
        for (void* p : s->owned) c->be->dfree(p);
    }
    delete s;
}

int pcv_shard_send_info(const pcv_shard_send* s, int* wide_records, int* digit_levels) {
    if (!s) return fail(PCV_ERR_INVALID, "null argument");
    if (wide_records) *wide_records = s->wide ? 1 : 0;
    if (digit_levels) *digit_levels = s->G0;
    return PCV_OK;
}

int pcv_shard_send_dest(const pcv_shard_send* s, const uint8_t** dev_dest, uint64_t* n) {
    if (!s || !dev_dest || !n) return fail(PCV_ERR_INVALID, "null argument");
    *dev_dest = s->dest;
    *n = s->n;
    return PCV_OK;
}

int pcv_shard_exchange_device(pcv_shard_send* sd, uint32_t k, const int32_t* cell_to_rank, uint32_t nranks, const uint64_t* dst_first, void* const* dst_rec,
                              void* const* dst_col, void* const* dst_dig, void* const* dst_intensity, uint64_t* rank_counts_out) {
    if (!sd || !cell_to_rank || !dst_first || !dst_rec || !dst_dig || !rank_counts_out) return fail(PCV_ERR_INVALID, "null argument");
    if (sd->wide && !dst_col) return fail(PCV_ERR_INVALID, "wide (Float64) records travel with a separate colour array: dst_col is required");
    if (nranks == 0 || nranks > (uint32_t)kMaxRanks) return fail(PCV_ERR_INVALID, "nranks must be 1..%d", kMaxRanks);
    if (k < 1 || (int)k > sd->G0) return fail(PCV_ERR_INVALID, "prefix levels must be 1..%d", sd->G0);
    if (sd->has_intensity && !dst_intensity) return fail(PCV_ERR_INVALID, "the points carry intensity: dst_intensity is required");
    API_TRY
    pcv_ctx* c = sd->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    for (uint32_t r = 0; r < nranks; ++r) rank_counts_out[r] = 0;
    if (sd->n == 0) return PCV_OK;
    ExchangeArgs a{};
    a.rec = sd->rec;
    a.col = sd->col;
    a.dig = sd->dig;
    a.intensity = sd->has_intensity ? sd->intensity : nullptr;
    a.tile_counts = sd->tile_counts;
    a.n = sd->n;
    a.ntiles = sd->ntiles;
    a.nbins = sd->nbins;
    a.cell_shift = 3 * (sd->G0 - (int)k);
    a.nranks = (int)nranks;
    a.wide = sd->wide;
    const int ncell = 1 << (3 * k);
    for (int i = 0; i < 64; ++i) a.cell_to_rank[i] = 0;
    for (int i = 0; i < ncell; ++i) {
        if (cell_to_rank[i] < 0 || cell_to_rank[i] >= (int32_t)nranks) return fail(PCV_ERR_INVALID, "cell_to_rank[%d] out of range", i);
        a.cell_to_rank[i] = (uint8_t)cell_to_rank[i];
    }
    for (uint32_t r = 0; r < nranks; ++r) {
        a.dst_first[r] = dst_first[r];
        a.dst_rec[r] = dst_rec[r];
        a.dst_col[r] = dst_col ? (uint32_t*)dst_col[r] : nullptr;
        a.dst_dig[r] = (uint8_t*)dst_dig[r];
        a.dst_intensity[r] = dst_intensity ? (float*)dst_intensity[r] : nullptr;
    }
    a.dest_out = sd->dest;
    const uint32_t grid = std::min<uint32_t>(sd->ntiles, (uint32_t)c->sm_count * 2u);
    if (sd->wide) {
        CU(cudaFuncSetAttribute(k_exchange_records<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ExSmem::bytes));
        k_exchange_records<true><<<grid, kExThreads, ExSmem::bytes, c->stream>>>(a);
    } else {
        CU(cudaFuncSetAttribute(k_exchange_records<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ExSmem::bytes));
        k_exchange_records<false><<<grid, kExThreads, ExSmem::bytes, c->stream>>>(a);
    }
    c->be->launches++;
    CU(cudaGetLastError());
    for (int d = 0; d < sd->nbins; ++d) rank_counts_out[a.cell_to_rank[d >> a.cell_shift]] += sd->bins[(size_t)d];
    CU(cudaStreamSynchronize(c->stream));  // every store of this rank has been issued and completed
    // the records are consumed: only the per-point destination (provenance look-ups) stays with the handle
    for (void*& p : sd->owned)
        if (p != (void*)sd->dest) {
            c->be->dfree(p);
            p = nullptr;
        }
    sd->rec = nullptr, sd->col = nullptr, sd->dig = nullptr, sd->tile_counts = nullptr;
    return PCV_OK;
    API_CATCH
}

// ---- fused exchange pass: the sender's FIRST PARTITION PASS writes into the owners' buffers --------------------
// Instead of moving the ingested records to their owner and partitioning them there, every sender runs the first two-level
// partition pass (root -> level-2 cells: rank by the carried digits, finish the level-2 codes, run the next pass's first step) on
// its own records and stores each bucket - a level-2 cell - straight into the buffers of the cell's owner over NVLink.  The layout
// on an owner is what its own first pass would have produced from (source rank, local index)-ordered input: cells in index order,
// inside a cell the senders in rank order.  All of it follows from the gathered per-sender histograms, so nothing is negotiated.
namespace {
struct FusedLayout {
    uint64_t slot_start[64];  // per cell: first slot (cell-major over ALL the owner's cells) on its owner
    uint64_t arena_off[64];   // leaf cells: first record in the owner's arena
    bool split[64];
    uint64_t total[64];
    uint64_t slots[kMaxRanks];
};
// nullptr on success, else why the fused pass cannot be used (the caller falls back to the exchange of ingested records)
const char* fused_layout(const LevelTable& lv, double resolution, uint64_t max_points, int nranks, const int32_t* c2r, const uint64_t* hist_all, FusedLayout& L) {
    if (lv.last_level < 2) return "the octree has fewer than two levels";
    for (int c = 0; c < 64; ++c) {
        L.total[c] = 0;
        for (int s = 0; s < nranks; ++s) L.total[c] += hist_all[(size_t)s * 64 + c];
    }
    for (int k1 = 0; k1 < 8; ++k1) {  // every non-empty level-1 node must split (distributed.py usable_prefix_levels guarantees it for k = 2)
        uint64_t t = 0;
        for (int k2 = 0; k2 < 8; ++k2) t += L.total[k1 * 8 + k2];
        if (t && !(t > max_points && lv.edge[1] > resolution)) return "a level-1 node does not split";
    }
    for (int r = 0; r < nranks; ++r) {
        uint64_t slot = 0, arena = 0;
        for (int c = 0; c < 64; ++c) {
            if (c2r[c] != r || !L.total[c]) continue;
            L.slot_start[c] = slot;
            L.split[c] = L.total[c] > max_points && lv.edge[2] > resolution;  // should_split_node on the global count (generation.rs:128-150)
            L.arena_off[c] = arena;
            if (!L.split[c]) arena += L.total[c];
            slot += L.total[c];
        }
        if (slot >= 0xFFFFFFFFull) return "an owner would receive 2^32 points or more";
        L.slots[r] = slot;
    }
    return nullptr;
}
}  // namespace

int pcv_shard_pass_device(pcv_shard_send* sd, uint32_t nranks, uint32_t rank, const int32_t* cell_to_rank, const uint64_t* hist_all, const pcv_shard_bufs* dst,
                          uint64_t* slots_out, uint64_t* first_bins_out) {
    if (!sd || !cell_to_rank || !hist_all || !dst) return fail(PCV_ERR_INVALID, "null argument");
    if (nranks == 0 || nranks > (uint32_t)kMaxRanks || rank >= nranks) return fail(PCV_ERR_INVALID, "nranks must be 1..%d and rank below it", kMaxRanks);
    if (sd->G0 != 2) return fail(PCV_ERR_UNSUPPORTED, "the fused exchange pass needs a two-level first pass");
    static_assert(sizeof(pcv_shard_bufs) == sizeof(RemoteBufs), "pcv_shard_bufs mirrors RemoteBufs");
    API_TRY
    pcv_ctx* c = sd->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < 64; ++i)
        if (cell_to_rank[i] < 0 || cell_to_rank[i] >= (int32_t)nranks) return fail(PCV_ERR_INVALID, "cell_to_rank[%d] out of range", i);
    for (int d = 0; d < 64; ++d)
        if (hist_all[(size_t)rank * 64 + d] != sd->bins[(size_t)d]) return fail(PCV_ERR_INVALID, "hist_all[rank] is not this handle's histogram");
    FusedLayout L{};
    const uint64_t max_points = c->cfg.max_points_per_node ? c->cfg.max_points_per_node : 100000;
    if (const char* why = fused_layout(sd->lv, sd->resolution, max_points, (int)nranks, cell_to_rank, hist_all, L)) return fail(PCV_ERR_UNSUPPORTED, "%s", why);
    if (slots_out)
        for (uint32_t r = 0; r < nranks; ++r) slots_out[r] = L.slots[r];
    if (first_bins_out)
        for (int cell = 0; cell < 64; ++cell) first_bins_out[cell] = cell_to_rank[cell] == (int32_t)rank ? L.total[cell] : 0;
    if (sd->n == 0) {
        sd->fused_done = true;
        return PCV_OK;
    }
    for (uint32_t r = 0; r < nranks; ++r) {
        if (!L.slots[r]) continue;
        if (!dst[r].rec_next || !dst[r].dig_next || !dst[r].arena || !dst[r].col_arena || (sd->wide && !dst[r].col_next) || (sd->has_intensity && !dst[r].intensity))
            return fail(PCV_ERR_INVALID, "destination buffers of rank %u are incomplete", r);
    }
    CudaBackend& be = *c->be;
    Scratch s(c);
    // this sender's bucket table: one bucket per non-empty local cell, in the owner's buffers behind the lower ranks' records
    std::vector<BucketDesc> bk(64, BucketDesc{});
    int nb = 0;
    for (int cell = 0; cell < 64; ++cell) {
        if (!sd->bins[(size_t)cell]) continue;
        uint64_t pre = 0;
        for (uint32_t r = 0; r < rank; ++r) pre += hist_all[(size_t)r * 64 + cell];
        BucketDesc b{};
        b.b0 = (uint16_t)cell;
        b.b1 = (uint16_t)(cell + 1);
        b.keep = 2;
        b.owner = (uint16_t)(cell_to_rank[cell] + 1);
        if (L.split[cell]) {
            b.kind = 0;
            b.dest = L.slot_start[cell] + pre;
        } else {
            b.kind = 1;
            b.dest = (L.arena_off[cell] + pre) | ((L.slot_start[cell] + pre) << 32);
        }
        bk[(size_t)nb++] = b;
    }
    std::vector<RemoteBufs> rb(nranks);
    for (uint32_t r = 0; r < nranks; ++r) std::memcpy(&rb[r], &dst[r], sizeof(RemoteBufs));
    BuildState hs{};
    hs.nnodes = 1;
    hs.pass[0].nactive = 1;
    hs.pass[0].ntiles = sd->ntiles;
    hs.pass[0].nchunks = (sd->ntiles + kChunkTiles - 1) / kChunkTiles;
    hs.pass[0].npoints = sd->n;
    ActiveDesc a0{};
    for (int a = 0; a < 3; ++a) a0.m[a] = sd->bmin[a];
    a0.e = sd->root_edge;
    a0.count = sd->n;
    a0.nchunks = hs.pass[0].nchunks;
    PassArgs pa{};
    pa.pass = 0;
    pa.level = 0;
    pa.G = 2;
    pa.nbins = 64;
    pa.Gn = std::min(2, sd->lv.last_level - 2);
    pa.wide = sd->wide;
    pa.remote = s.upload(rb.data(), rb.size());
    pa.int_in = sd->has_intensity ? sd->intensity : nullptr;
    pa.rec_in = sd->rec;
    pa.col_in = sd->col;
    pa.dig_in = sd->dig;
    pa.st = s.upload(&hs, 1);
    pa.active = s.upload(&a0, 1);
    pa.tile_active = sd->tile_active;
    pa.tile_counts = sd->tile_counts;
    pa.buckets = s.upload(bk.data(), bk.size());
    pa.cap_active = 1;
    pa.cap_tiles = sd->ntiles;
    pa.resolution = sd->resolution;
    pa.lv = sd->lv;
    set_level_constants(pa, sd->lv);
    be.partition(pa);
    CU(cudaStreamSynchronize(c->stream));  // every store of this rank has been issued and completed
    // the records are consumed; the digits - the level-2 cell of every local point - stay for provenance look-ups
    for (void*& p : sd->owned)
        if (p != (void*)sd->dig) {
            be.dfree(p);
            p = nullptr;
        }
    sd->rec = nullptr, sd->col = nullptr, sd->tile_counts = nullptr, sd->tile_active = nullptr, sd->dest = nullptr;
    sd->fused_done = true;
    return PCV_OK;
    API_CATCH
}

int pcv_shard_send_cells(const pcv_shard_send* s, const uint8_t** dev_cells, uint64_t* n) {
    if (!s || !dev_cells || !n) return fail(PCV_ERR_INVALID, "null argument");
    if (!s->fused_done) return fail(PCV_ERR_INVALID, "the handle did not run pcv_shard_pass_device");
    *dev_cells = s->dig;
    *n = s->n;
    return PCV_OK;
}

int pcv_build_octree_after_pass_device(pcv_ctx* c, const pcv_shard_bufs* own, uint64_t nslots, const uint64_t* first_bins, double resolution,
                                       const double bmin_in[3], const double bmax_in[3], const uint64_t* prefix_counts, pcv_octree** out) {
    if (!c || !own || !first_bins || !bmin_in || !bmax_in || !prefix_counts || !out) return fail(PCV_ERR_INVALID, "null argument");
    if (nslots && (!own->rec_next || !own->dig_next || !own->arena || !own->col_arena)) return fail(PCV_ERR_INVALID, "incomplete buffers");
    *out = nullptr;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ShardSpec sp;
    sp.k = 2;
    sp.counts = prefix_counts;
    ExternalRecords ext;
    ext.rec = own->rec_next;
    ext.col = (uint32_t*)own->col_next;
    ext.dig = (uint8_t*)own->dig_next;
    ext.n = nslots;
    ext.present = true;
    ext.col_in_record = own->col_next == nullptr;
    ext.after_first_pass = true;
    ext.arena = own->arena;
    ext.col_arena = (uint32_t*)own->col_arena;
    ext.first_bins = first_bins;
    PointsView v{};
    v.stride = 1;
    v.n = nslots;
    v.intensity = (const float*)own->intensity;
    v.rgb = reinterpret_cast<const uint8_t*>(own->rec_next);  // non-null marker: colours travel with the records
    return build_impl(c, v, resolution, bmin_in, bmax_in, out, &sp, &ext);
    API_CATCH
}

int pcv_build_octree_from_records_device(pcv_ctx* c, void* rec, uint32_t* col, uint8_t* dig, const float* intensity, uint64_t n, double resolution,
                                         const double bmin_in[3], const double bmax_in[3], uint32_t k, const uint64_t* prefix_counts, pcv_octree** out) {
    if (!c || !bmin_in || !bmax_in || !prefix_counts || !out || (n && (!rec || !dig))) return fail(PCV_ERR_INVALID, "null argument");
    if (k < 1 || k > 3) return fail(PCV_ERR_INVALID, "prefix levels k must be 1..3");
    *out = nullptr;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ShardSpec sp;
    sp.k = (int)k;
    sp.counts = prefix_counts;
    ExternalRecords ext;
    ext.rec = rec;
    ext.col = col;
    ext.dig = dig;
    ext.n = n;
    ext.present = true;
    ext.col_in_record = col == nullptr;  // narrow records as exchanged: {code x 3, colour}
    PointsView v{};
    v.stride = 1;
    v.n = n;
    v.intensity = intensity;
    v.rgb = reinterpret_cast<const uint8_t*>(rec);  // non-null marker: colours travel with the records
    return build_impl(c, v, resolution, bmin_in, bmax_in, out, &sp, &ext);
    API_CATCH
}

int pcv_octree_node_nsub(const pcv_octree* o, uint64_t hi, uint64_t lo, uint64_t* nsub_out) {
    if (!o || !nsub_out) return fail(PCV_ERR_INVALID, "null argument");
    const int i = o->find(hi, lo);
    if (i < 0 || (size_t)i >= o->nsub.size()) return fail(PCV_ERR_NOT_FOUND, "node %s not found", node_name(hi, lo).c_str());
    *nsub_out = o->nsub[(size_t)i];
    return PCV_OK;
}

int pcv_octree_nsub_all(const pcv_octree* o, uint64_t* out, uint64_t cap) {
    if (!o || (!out && cap)) return fail(PCV_ERR_INVALID, "null argument");
    if (cap < o->nsub.size()) return fail(PCV_ERR_INVALID, "capacity too small");
    if (!o->nsub.empty()) memcpy(out, o->nsub.data(), o->nsub.size() * 8);
    return PCV_OK;
}

int pcv_assemble_top(pcv_ctx* c, double resolution, const double bmin_in[3], const double bmax_in[3], uint32_t k, const uint64_t* prefix_counts,
                     const uint64_t* unit_nsub, const void* xyz_codes, const uint8_t* rgb, const float* intensity, uint64_t npoints, pcv_octree** out) {
    if (!c || !bmin_in || !bmax_in || !prefix_counts || !unit_nsub || !out || (npoints && (!xyz_codes || !rgb)))
        return fail(PCV_ERR_INVALID, "null argument");
    *out = nullptr;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; ++a) {
        bmin[a] = std::fmin(bmin_in[a], bmax_in[a]);
        bmax[a] = std::fmax(bmin_in[a], bmax_in[a]);
    }
    BuildResult R = assemble_top(*c->be, resolution, bmin, bmax, (int)k, prefix_counts, unit_nsub, (const uint8_t*)xyz_codes, rgb, intensity, npoints);
    CU(cudaStreamSynchronize(c->stream));
    *out = octree_from_result(c, R, resolution, bmin, bmax, intensity != nullptr);
    return PCV_OK;
    API_CATCH
}

}  // extern "C"
