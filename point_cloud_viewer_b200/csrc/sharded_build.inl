// sharded_build.inl — the multi-GPU build_octree behind the C ABI (included at the end of pcv_api.cu): one call per rank,
// the collectives come from the caller as three callbacks (NCCL, MPI, torch.distributed ... - anything with an all-reduce, an
// all-gather and a barrier), so the path is reachable from the Rust shim of INTEGRATION.md without any Python.
// Steps: ingest the local points -> all-reduce the level-k cell histogram -> cells to ranks ->
// one kernel stores every record into its owner's slab over NVLink (CUDA-IPC mapping, set up once and cached per context) ->
// every owner builds from its slab -> the nodes above level k are assembled on rank 0 from the collectors' content.

namespace {

using namespace pcv;

struct ShardSlab {
    void* ptr = nullptr;
    uint64_t cap = 0;
    int world = 0;
    bool wide = false, with_int = false;
    std::vector<void*> peer;
    // one allocation per rank: next-pass records | their colours (wide records only) | intensities | digits | leaf arena | its colours
    uint64_t off_col = 0, off_int = 0, off_dig = 0, off_arena = 0, off_acol = 0, bytes = 0;
    void* rec(int r) const { return peer[(size_t)r]; }
    void* col(int r) const { return (uint8_t*)peer[(size_t)r] + off_col; }
    void* inten(int r) const { return (uint8_t*)peer[(size_t)r] + off_int; }
    void* dig(int r) const { return (uint8_t*)peer[(size_t)r] + off_dig; }
    pcv_shard_bufs bufs(int r) const {
        uint8_t* b = (uint8_t*)peer[(size_t)r];
        return pcv_shard_bufs{b, wide ? b + off_col : nullptr, b + off_dig, b + off_arena, b + off_acol, with_int ? b + off_int : nullptr};
    }
};
std::mutex g_slab_mu;
std::map<pcv_ctx*, ShardSlab> g_slabs;
// wall-clock phases of the last pcv_build_octree_sharded per context (every phase ends in a stream synchronisation or a barrier):
// ingest + histogram, all-reduce + plan, exchange, local build, top assembly; [5] = 1 if the fused exchange pass ran
std::map<pcv_ctx*, std::array<double, 6>> g_phases;

#define COMM(x)                                                                                    \
    do {                                                                                           \
        if ((x) != 0) throw BuildError(PCV_ERR_INVALID, "communication callback failed: " #x);    \
    } while (0)
#define PCVX(x)                                                                        \
    do {                                                                               \
        const int rc_ = (x);                                                           \
        if (rc_ != PCV_OK) throw BuildError(rc_, std::string(pcv_last_error()));      \
    } while (0)

// distributed.py: level_counts / usable_prefix_levels / assign_cells, restated
std::vector<uint64_t> level_counts_of(const std::vector<uint64_t>& ck, int k, int level) {
    std::vector<uint64_t> out((size_t)1 << (3 * level), 0);
    const int shift = 3 * (k - level);
    for (size_t i = 0; i < ck.size(); ++i) out[i >> shift] += ck[i];
    return out;
}
int usable_prefix_levels(const std::vector<uint64_t>& ck, int k, double root_edge, double resolution, uint64_t max_points) {
    int ok = 1;
    double edge = root_edge;
    for (int j = 1; j < k; ++j) {
        edge = edge / 2.0;
        const std::vector<uint64_t> c = level_counts_of(ck, k, j);
        bool any = false, all = true;
        for (uint64_t v : c)
            if (v) {
                any = true;
                all = all && v > max_points;
            }
        if (any && all && edge > resolution)
            ok = j + 1;
        else
            break;
    }
    return ok;
}
std::vector<int32_t> assign_cells_lpt(const std::vector<uint64_t>& counts, int nranks) {
    std::vector<size_t> order(counts.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return counts[a] != counts[b] ? counts[a] > counts[b] : a < b; });
    std::vector<uint64_t> load((size_t)nranks, 0);
    std::vector<int32_t> out(counts.size(), 0);
    for (size_t c : order) {
        if (counts[c] == 0) continue;
        int r = 0;
        for (int i = 1; i < nranks; ++i)
            if (load[(size_t)i] < load[(size_t)r]) r = i;
        out[c] = r;
        load[(size_t)r] += counts[c];
    }
    return out;
}

ShardSlab& slab_for(pcv_ctx* c, const pcv_comm* comm, uint64_t need, bool wide, bool with_int) {
    std::lock_guard<std::mutex> g(g_slab_mu);
    ShardSlab& s = g_slabs[c];
    if (s.ptr && s.cap >= need && s.world == comm->world && s.wide == wide && s.with_int == with_int) return s;
    if (s.ptr) {  // collective teardown: nobody may still write into a slab that is about to disappear
        COMM(comm->barrier(comm->user));
        for (int r = 0; r < s.world; ++r)
            if (r != comm->rank && s.peer[(size_t)r]) pcv_ipc_close(c, s.peer[(size_t)r]);
        COMM(comm->barrier(comm->user));
        pcv_ipc_free(c, s.ptr);
        s = ShardSlab();
    }
    s.cap = (((uint64_t)((double)need * 1.05) + 4096 + 4095) / 4096) * 4096;
    s.world = comm->world;
    s.wide = wide;
    s.with_int = with_int;
    const uint64_t rs = wide ? 32 : 16;
    s.off_col = rs * s.cap + 256;
    s.off_int = s.off_col + 4 * s.cap + 256;
    s.off_dig = s.off_int + (with_int ? 4 * s.cap + 256 : 0);
    s.off_arena = s.off_dig + ((s.cap + 256 + 255) / 256) * 256;
    s.off_acol = s.off_arena + rs * s.cap + 256;
    s.bytes = s.off_acol + 4 * s.cap + 256;
    uint8_t handle[64];
    PCVX(pcv_ipc_alloc(c, s.bytes, &s.ptr, handle));
    std::vector<uint8_t> all((size_t)comm->world * 64);
    COMM(comm->allgather(comm->user, handle, 64, all.data()));
    s.peer.assign((size_t)comm->world, nullptr);
    for (int r = 0; r < comm->world; ++r) {
        if (r == comm->rank)
            s.peer[(size_t)r] = s.ptr;
        else
            PCVX(pcv_ipc_open(c, all.data() + (size_t)r * 64, &s.peer[(size_t)r]));
    }
    COMM(comm->barrier(comm->user));
    return s;
}

}  // namespace

static void sharded_forget(pcv_ctx* c) {
    std::lock_guard<std::mutex> g(g_slab_mu);
    auto it = g_slabs.find(c);
    if (it == g_slabs.end()) return;
    ShardSlab& s = it->second;
    for (int r = 0; r < s.world; ++r)
        if (s.peer[(size_t)r] && s.peer[(size_t)r] != s.ptr) cudaIpcCloseMemHandle(s.peer[(size_t)r]);
    if (s.ptr) cudaFree(s.ptr);
    g_slabs.erase(it);
    g_phases.erase(c);
}

extern "C" {

int pcv_build_octree_sharded(pcv_ctx* c, const pcv_comm* comm, const pcv_points* dp, double resolution, const double bmin_in[3], const double bmax_in[3],
                             uint32_t prefix_levels, pcv_octree** local_out, pcv_octree** top_out, uint32_t* k_out, int32_t* cell_to_rank_out, uint64_t* unit_nsub_out,
                             uint64_t* recv_points_out, pcv_shard_send** send_out) {
    if (!c || !comm || !dp || !bmin_in || !bmax_in || !local_out || !top_out) return fail(PCV_ERR_INVALID, "null argument");
    if (!comm->allreduce_sum_u64 || !comm->allgather || !comm->barrier || comm->world < 1 || comm->rank < 0 || comm->rank >= comm->world || comm->world > kMaxRanks)
        return fail(PCV_ERR_INVALID, "invalid communicator");
    if (prefix_levels < 1 || prefix_levels > 2) return fail(PCV_ERR_INVALID, "prefix levels must be 1 or 2");
    *local_out = nullptr;
    *top_out = nullptr;
    pcv_shard_send* send = nullptr;
    API_TRY
    const int R = comm->world, me = comm->rank;
    std::array<double, 6> ph{};
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](int i) {
        const auto t = std::chrono::steady_clock::now();
        ph[(size_t)i] += std::chrono::duration<double, std::milli>(t - t_last).count();
        t_last = t;
    };
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; ++a) {
        bmin[a] = std::fmin(bmin_in[a], bmax_in[a]);
        bmax[a] = std::fmax(bmin_in[a], bmax_in[a]);
    }
    const double E = std::fmax(std::fmax(bmax[0] - bmin[0], bmax[1] - bmin[1]), bmax[2] - bmin[2]);
    int k = (int)prefix_levels;
    // (1) local ingest + histogram, global histogram
    std::vector<uint64_t> local_hist((size_t)1 << (3 * k), 0);
    PCVX(pcv_shard_ingest_device(c, dp, resolution, bmin, bmax, (uint32_t)k, local_hist.data(), &send));
    mark(0);
    // one all-gather of the per-rank histograms carries everything the plan needs: the global counts (their sum), the count
    // matrix (who sends how much to whom) and the per-sender cell counts of the fused exchange pass
    std::vector<uint64_t> hist_all((size_t)R * local_hist.size());
    COMM(comm->allgather(comm->user, local_hist.data(), (uint64_t)local_hist.size() * 8, hist_all.data()));
    std::vector<uint64_t> counts_k(local_hist.size(), 0);
    for (int s = 0; s < R; ++s)
        for (size_t cell = 0; cell < counts_k.size(); ++cell) counts_k[cell] += hist_all[(size_t)s * counts_k.size() + cell];
    const int k2 = usable_prefix_levels(counts_k, k, E, resolution, c->cfg.max_points_per_node);
    if (k2 < k) {
        std::vector<uint64_t> ha;
        for (int s = 0; s < R; ++s) {
            const std::vector<uint64_t> row(hist_all.begin() + (size_t)s * counts_k.size(), hist_all.begin() + (size_t)(s + 1) * counts_k.size());
            const std::vector<uint64_t> red = level_counts_of(row, k, k2);
            ha.insert(ha.end(), red.begin(), red.end());
        }
        hist_all.swap(ha);
        counts_k = level_counts_of(counts_k, k, k2);
        local_hist = level_counts_of(local_hist, k, k2);
        k = k2;
    }
    std::vector<uint64_t> prefix_counts;
    for (int j = 1; j <= k; ++j) {
        const std::vector<uint64_t> lc = level_counts_of(counts_k, k, j);
        prefix_counts.insert(prefix_counts.end(), lc.begin(), lc.end());
    }
    // (2) cells -> ranks, count matrix
    const std::vector<int32_t> c2r = assign_cells_lpt(counts_k, R);
    std::vector<uint64_t> M((size_t)R * R, 0);  // M[s * R + d]
    for (int s = 0; s < R; ++s)
        for (size_t cell = 0; cell < counts_k.size(); ++cell) M[(size_t)s * R + (size_t)c2r[cell]] += hist_all[(size_t)s * counts_k.size() + cell];
    uint64_t need = 0, n_recv = 0;
    for (int d = 0; d < R; ++d) {
        uint64_t t = 0;
        for (int s = 0; s < R; ++s) t += M[(size_t)s * R + d];
        need = std::max(need, t);
        if (d == me) n_recv = t;
    }
    int wide = 0;
    PCVX(pcv_shard_send_info(send, &wide, nullptr));
    const bool with_int = dp->intensity != nullptr;
    ShardSlab& slab = slab_for(c, comm, need, wide != 0, with_int);
    mark(1);
    // (3a) fused exchange pass: every sender's first partition pass stores into the owners' buffers; the owner's build starts at
    // its second pass.  Falls back to (3b) when the layout does not allow it (same decision on every rank: same inputs).
    pcv_octree* local = nullptr;
    // Measured on 8 x B200 (1e9 points per GPU): with 56 of a sender's 64 buckets in peer memory the fused pass's short remote runs
    // (28 records per bucket and tile) collapse the NVLink store rate (315 ms against 22 ms on two GPUs), while the exchange of
    // ingested records - 7 long runs per tile - holds 515 GB/s per GPU (28.9 ms).  So the fused pass is the default up to two
    // ranks; PCV_FUSED_PASS=1 forces it, PCV_NO_FUSED_PASS=1 disables it.
    bool fused = k == 2 && !std::getenv("PCV_NO_FUSED_PASS") && (R <= 2 || std::getenv("PCV_FUSED_PASS"));
    if (fused) {
        std::vector<uint64_t> first_bins(64, 0), slots((size_t)R, 0);
        std::vector<pcv_shard_bufs> dst((size_t)R);
        for (int d = 0; d < R; ++d) dst[(size_t)d] = slab.bufs(d);
        COMM(comm->barrier(comm->user));  // no peer is still building out of its slab
        const int rc = pcv_shard_pass_device(send, (uint32_t)R, (uint32_t)me, c2r.data(), hist_all.data(), dst.data(), slots.data(), first_bins.data());
        if (rc == PCV_ERR_UNSUPPORTED) {
            fused = false;
        } else {
            PCVX(rc);
            if (slots[(size_t)me] != n_recv) throw BuildError(PCV_ERR_CUDA, "internal: slot count and count matrix disagree");
            COMM(comm->barrier(comm->user));  // every rank's stores have landed
            mark(2);
            ph[5] = 1;
            const pcv_shard_bufs own = slab.bufs(me);
            PCVX(pcv_build_octree_after_pass_device(c, &own, n_recv, first_bins.data(), resolution, bmin, bmax, prefix_counts.data(), &local));
        }
    }
    if (!fused) {
    // (3b) one kernel: every ingested record into its owner's slab
    std::vector<uint64_t> first((size_t)R, 0), got((size_t)R, 0);
    std::vector<void*> drec((size_t)R), dcol((size_t)R), ddig((size_t)R), dint((size_t)R);
    for (int d = 0; d < R; ++d) {
        for (int s = 0; s < me; ++s) first[(size_t)d] += M[(size_t)s * R + d];
        drec[(size_t)d] = slab.rec(d);
        dcol[(size_t)d] = slab.col(d);
        ddig[(size_t)d] = slab.dig(d);
        dint[(size_t)d] = slab.inten(d);
    }
    COMM(comm->barrier(comm->user));  // no peer is still using its slab as build scratch
    PCVX(pcv_shard_exchange_device(send, (uint32_t)k, c2r.data(), (uint32_t)R, first.data(), drec.data(), wide ? dcol.data() : nullptr, ddig.data(),
                                   with_int ? dint.data() : nullptr, got.data()));
    for (int d = 0; d < R; ++d)
        if (got[(size_t)d] != M[(size_t)me * R + d]) throw BuildError(PCV_ERR_CUDA, "internal: local histogram and exchange disagree");
    COMM(comm->barrier(comm->user));  // every rank's stores have landed
    mark(2);
    // (4) the owner's build
    PCVX(pcv_build_octree_from_records_device(c, n_recv ? slab.rec(me) : nullptr, (wide && n_recv) ? (uint32_t*)slab.col(me) : nullptr, n_recv ? (uint8_t*)slab.dig(me) : nullptr,
                                              (with_int && n_recv) ? (const float*)slab.inten(me) : nullptr, n_recv, resolution, bmin, bmax, (uint32_t)k,
                                              prefix_counts.data(), &local));
    }
    *local_out = local;
    mark(3);
    // (5) top of the tree: unit sizes, collectors' content -> rank 0
    const size_t ncell = (size_t)1 << (3 * k);
    std::vector<uint64_t> unit_nsub(ncell, 0);
    const uint64_t idx_mask = ((uint64_t)1 << 60) - 1;
    for (size_t i = 0; i < local->nodes.size(); ++i)
        if (local->nodes[i].level == k) unit_nsub[(size_t)(local->nodes[i].id_low & idx_mask)] = local->nsub[i];
    COMM(comm->allreduce_sum_u64(comm->user, unit_nsub.data(), ncell));
    const LevelTable lv = make_level_table(E, resolution, bmin);
    const uint64_t bpc = (uint64_t)enc_bytes(lv.enc[k - 1]);
    const uint64_t per_pt = 3 * bpc + 3 + (with_int ? 4 : 0);
    std::vector<std::vector<std::pair<uint64_t, uint64_t>>> owned((size_t)R);  // per rank: (cell, count) in cell order
    for (size_t cell = 0; cell < ncell; ++cell)
        if (unit_nsub[cell]) owned[(size_t)c2r[cell]].push_back({cell, (unit_nsub[cell] + 7) / 8});
    uint64_t maxbytes = 16;
    std::vector<uint64_t> tot((size_t)R, 0);
    for (int r = 0; r < R; ++r) {
        for (auto& p : owned[(size_t)r]) tot[(size_t)r] += p.second;
        maxbytes = std::max(maxbytes, tot[(size_t)r] * per_pt);
    }
    std::vector<uint8_t> mine((size_t)maxbytes, 0);
    {
        // this rank's collectors (level k-1 nodes with points): their content is the concatenation, in child order, of what the
        // local children gave up
        const uint64_t T = tot[(size_t)me];
        uint64_t ox = 0, orr = T * 3 * bpc, oi = T * (3 * bpc + 3);
        std::lock_guard<std::mutex> g(c->mu);
        CU(cudaSetDevice(c->device));
        for (size_t i = 0; i < local->nodes.size(); ++i) {
            const pcv_node_meta& m = local->nodes[i];
            if (m.level != k - 1 || m.num_points == 0) continue;
            const uint64_t n = (uint64_t)m.num_points, pidx = k > 1 ? (m.id_low & idx_mask) : 0;
            uint64_t expect = 0;
            for (int ch = 0; ch < 8; ++ch) {
                const size_t cell = (size_t)(pidx * 8 + (uint64_t)ch);
                if (c2r[cell] == me && unit_nsub[cell]) expect += (unit_nsub[cell] + 7) / 8;
            }
            if (expect != n) throw BuildError(PCV_ERR_CUDA, "internal: collector size does not match the unit sizes");
            // cells of one collector are consecutive in this rank's cell-ordered buffer
            CU(cudaMemcpyAsync(mine.data() + ox, local->d_xyz + m.xyz_byte_offset, n * 3 * bpc, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaMemcpyAsync(mine.data() + orr, local->d_rgb + 3 * m.point_offset, n * 3, cudaMemcpyDeviceToHost, c->stream));
            if (with_int) CU(cudaMemcpyAsync(mine.data() + oi, local->d_intensity + m.point_offset, n * 4, cudaMemcpyDeviceToHost, c->stream));
            ox += n * 3 * bpc;
            orr += n * 3;
            oi += n * 4;
        }
        CU(cudaStreamSynchronize(c->stream));
        if (ox != T * 3 * bpc) throw BuildError(PCV_ERR_CUDA, "internal: collectors do not cover this rank's cells");
    }
    std::vector<uint8_t> all((size_t)R * maxbytes);
    COMM(comm->allgather(comm->user, mine.data(), maxbytes, all.data()));
    if (me == 0) {
        struct Piece {
            const uint8_t *x, *r, *i;
            uint64_t cnt;
        };
        std::vector<Piece> pieces(ncell, Piece{nullptr, nullptr, nullptr, 0});
        uint64_t npts = 0;
        for (int r = 0; r < R; ++r) {
            const uint8_t* row = all.data() + (size_t)r * maxbytes;
            uint64_t ox = 0, orr = tot[(size_t)r] * 3 * bpc, oi = tot[(size_t)r] * (3 * bpc + 3);
            for (auto& p : owned[(size_t)r]) {
                pieces[(size_t)p.first] = Piece{row + ox, row + orr, row + oi, p.second};
                ox += p.second * 3 * bpc;
                orr += p.second * 3;
                oi += p.second * 4;
                npts += p.second;
            }
        }
        std::vector<uint8_t> tx((size_t)(npts * 3 * bpc)), tr((size_t)(npts * 3));
        std::vector<float> tin(with_int ? (size_t)npts : 0);
        uint64_t o = 0;
        for (size_t cell = 0; cell < ncell; ++cell) {
            const Piece& p = pieces[cell];
            if (!p.cnt) continue;
            std::memcpy(tx.data() + o * 3 * bpc, p.x, (size_t)(p.cnt * 3 * bpc));
            std::memcpy(tr.data() + o * 3, p.r, (size_t)(p.cnt * 3));
            if (with_int) std::memcpy(tin.data() + o, p.i, (size_t)(p.cnt * 4));
            o += p.cnt;
        }
        PCVX(pcv_assemble_top(c, resolution, bmin, bmax, (uint32_t)k, prefix_counts.data(), unit_nsub.data(), tx.data(), tr.data(), with_int ? tin.data() : nullptr, npts,
                              top_out));
    }
    mark(4);
    {
        std::lock_guard<std::mutex> g(g_slab_mu);
        g_phases[c] = ph;
    }
    if (const char* tv = std::getenv("PCV_TIMING"); tv && (me == 0 || tv[0] == '2'))
        fprintf(stderr, "[pcv sharded C r%d] ingest + histogram %.1f  all-reduce + plan %.1f  exchange%s %.1f  local build %.1f  top assembly %.1f ms\n", me, ph[0], ph[1],
                ph[5] != 0 ? " (fused pass)" : "", ph[2], ph[3], ph[4]);
    if (k_out) *k_out = (uint32_t)k;
    if (recv_points_out) *recv_points_out = n_recv;
    if (cell_to_rank_out) std::memcpy(cell_to_rank_out, c2r.data(), ncell * sizeof(int32_t));
    if (unit_nsub_out) std::memcpy(unit_nsub_out, unit_nsub.data(), ncell * sizeof(uint64_t));
    if (send_out)
        *send_out = send;
    else
        pcv_shard_send_free(send);
    return PCV_OK;
    }
    catch (const BuildError& e) {
        if (send) pcv_shard_send_free(send);
        return fail(e.code, "%s", e.what());
    }
    catch (const std::exception& e) {
        if (send) pcv_shard_send_free(send);
        return fail(PCV_ERR_INVALID, "%s", e.what());
    }
}

int pcv_sharded_phases(pcv_ctx* c, double out[6]) {
    if (!c || !out) return fail(PCV_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> g(g_slab_mu);
    auto it = g_phases.find(c);
    if (it == g_phases.end()) return fail(PCV_ERR_NOT_FOUND, "no sharded build has run on this context");
    for (int i = 0; i < 6; ++i) out[i] = it->second[(size_t)i];
    return PCV_OK;
}

int pcv_sharded_release(pcv_ctx* c, const pcv_comm* comm) {
    if (!c || !comm) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    std::lock_guard<std::mutex> g(g_slab_mu);
    auto it = g_slabs.find(c);
    if (it == g_slabs.end() || !it->second.ptr) return PCV_OK;
    ShardSlab& s = it->second;
    COMM(comm->barrier(comm->user));
    for (int r = 0; r < s.world; ++r)
        if (r != comm->rank && s.peer[(size_t)r]) pcv_ipc_close(c, s.peer[(size_t)r]);
    COMM(comm->barrier(comm->user));
    pcv_ipc_free(c, s.ptr);
    g_slabs.erase(it);
    return PCV_OK;
    API_CATCH
}

}  // extern "C"
