// ply_api.inl — C ABI of the PLY input path (included by pcv_api.cu inside extern "C").

static PlyUnpackArgs ply_args(const pcv_ply_info& info, double* x, double* y, double* z, uint8_t* rgb, float* intensity) {
    PlyUnpackArgs a{};
    a.record_bytes = info.record_bytes;
    a.tile_points = ply_tile_points(info.record_bytes);
    for (int k = 0; k < 3; ++k) {
        a.type[k] = info.type_xyz[k];
        a.off[k] = info.off_xyz[k];
        a.off_rgb[k] = info.off_rgb[k];
        a.offset[k] = info.offset[k];
    }
    a.off_intensity = info.off_intensity;
    a.has_color = info.has_color;
    a.has_intensity = info.has_intensity;
    a.x = x, a.y = y, a.z = z, a.rgb = rgb, a.intensity = intensity;
    return a;
}

// Launches the unpack kernel over `n` records at `raw` (device) writing outputs from index `out_first`; the blocks'
// bounding boxes go to partial[tile_first .. ).
static void ply_launch(pcv_ctx* c, PlyUnpackArgs a, const uint8_t* raw, uint64_t n, uint64_t out_first, double* partial) {
    if (n == 0) return;
    a.raw = raw;
    a.n = n;
    a.out_first = out_first;
    a.partial = partial;
    const size_t sm = ply_smem_bytes(a.record_bytes, a.tile_points);
    // the dynamic shared memory opt-in of k_ply_unpack is per device: done in pcv_create for the context's GPU
    const uint32_t blocks = (uint32_t)((n + a.tile_points - 1) / a.tile_points);
    c->be->prof_begin(CudaBackend::K_PLY, n * ((uint64_t)a.record_bytes + 24 + (a.has_color && a.rgb ? 3 : 0) + (a.has_intensity && a.intensity ? 4 : 0)));
    k_ply_unpack<<<blocks, kPlyThreads, sm, c->stream>>>(a);
    c->be->prof_end();
    ++c->be->launches;
    CU(cudaGetLastError());
}

static void ply_reduce_bbox(pcv_ctx* c, const double* d_partial, uint64_t ntiles, double bmin[3], double bmax[3]) {
    for (int k = 0; k < 3; ++k) bmin[k] = bmax[k] = 0.0;  // Aabb::zero for an empty file (generation.rs:269)
    if (ntiles == 0) return;
    std::vector<double> h((size_t)ntiles * 6);
    c->be->d2h(h.data(), d_partial, h.size() * 8);
    for (int k = 0; k < 3; ++k) {
        bmin[k] = h[k];
        bmax[k] = h[3 + k];
    }
    for (uint64_t t = 1; t < ntiles; ++t)
        for (int k = 0; k < 3; ++k) {
            bmin[k] = std::fmin(bmin[k], h[(size_t)t * 6 + k]);
            bmax[k] = std::fmax(bmax[k], h[(size_t)t * 6 + 3 + k]);
        }
}

// File body -> device SoA arrays + bounding box.
// A pool of reader threads copies the body piece by piece (4 MiB preads: page cache / NVMe queue depth and the copy into
// pinned memory both scale with threads) into a ring of pinned chunks; the calling thread forwards every completed chunk
// to the GPU (H2D copy + unpack kernel on the context's stream) and hands the slot back once its copy has left pinned
// memory.  Reading chunk k + 1 and k + 2 overlaps the transfer and unpacking of chunk k.
static void ply_load(pcv_ctx* c, const char* path, const pcv_ply_info& info, double* x, double* y, double* z, uint8_t* rgb, float* intensity,
                     double bmin[3], double bmax[3]) {
    ply_validate(info);
    const uint64_t n = info.num_points;
    PlyUnpackArgs a = ply_args(info, x, y, z, rgb, intensity);
    const uint64_t T = a.tile_points;
    const uint64_t ntiles = (n + T - 1) / T;
    if (n == 0) {
        ply_reduce_bbox(c, nullptr, 0, bmin, bmax);
        return;
    }
    PlyFile f(path);
    if (f.fd < 0) throw BuildError(PCV_ERR_IO, "Could not open input file.");
    struct stat st;
    if (fstat(f.fd, &st) != 0) throw BuildError(PCV_ERR_IO, "stat failed");
    if ((uint64_t)st.st_size < info.header_bytes + n * info.record_bytes) throw BuildError(PCV_ERR_IO, "truncated PLY body");
    const uint64_t chunk_points = std::max<uint64_t>(T, (((uint64_t)64 << 20) / info.record_bytes) / T * T);
    const uint64_t chunk_bytes = chunk_points * info.record_bytes;
    const uint64_t nchunks = (n + chunk_points - 1) / chunk_points;
    const uint64_t piece = (uint64_t)4 << 20;
    const uint64_t ppc = (chunk_bytes + piece - 1) / piece;  // pieces per (full) chunk
    constexpr int kPin = 3, kDev = 2;
    CudaBackend& be = *c->be;
    uint8_t** pin = c->ply_pin;  // pinned staging ring, kept by the context between calls (allocating it costs ~50 ms)
    uint8_t* dev[kDev] = {nullptr, nullptr};
    cudaEvent_t copied[kPin] = {nullptr, nullptr, nullptr};
    double* d_partial = (double*)be.dmalloc(ntiles * 6 * 8);
    const bool timing = std::getenv("PCV_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };

    // reader pool state
    std::mutex mu;
    std::condition_variable cv_slot, cv_done;
    uint64_t released = kPin;  // chunks [0, released) may be written into their slots
    std::vector<uint32_t> done(nchunks, 0);
    std::atomic<uint64_t> next{0};
    int err = 0;
    bool stop = false;
    auto chunk_len = [&](uint64_t k) { return std::min(chunk_points, n - k * chunk_points) * info.record_bytes; };
    auto chunk_pieces = [&](uint64_t k) { return (uint32_t)((chunk_len(k) + piece - 1) / piece); };
    auto reader = [&]() {
        for (;;) {
            const uint64_t p = next.fetch_add(1);
            const uint64_t k = p / ppc, q = p % ppc;
            if (k >= nchunks) return;
            if (q >= chunk_pieces(k)) continue;  // the last chunk is shorter
            {
                std::unique_lock<std::mutex> l(mu);
                cv_slot.wait(l, [&] { return stop || k < released; });
                if (stop) return;
            }
            const uint64_t o = q * piece, len = std::min(piece, chunk_len(k) - o);
            uint8_t* dst = pin[k % kPin] + o;
            const uint64_t fo = info.header_bytes + k * chunk_bytes + o;
            uint64_t got_total = 0;
            int e = 0;
            while (got_total < len) {
                const ssize_t got = ::pread(f.fd, dst + got_total, len - got_total, (off_t)(fo + got_total));
                if (got <= 0) {
                    e = got == 0 ? 1 : 2;
                    break;
                }
                got_total += (uint64_t)got;
            }
            std::lock_guard<std::mutex> l(mu);
            if (e) err = e;
            ++done[k];
            cv_done.notify_all();
        }
    };
    std::vector<std::thread> pool;
    auto shutdown = [&]() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cv_slot.notify_all();
        for (auto& t : pool) t.join();
        pool.clear();
    };
    auto cleanup = [&]() {
        shutdown();
        cudaStreamSynchronize(c->stream);
        for (auto& d : dev) be.dfree(d);
        for (auto& e : copied)
            if (e) cudaEventDestroy(e);
        be.dfree(d_partial);
    };
    double ms_wait_read = 0, ms_wait_copy = 0;
    try {
        if (c->ply_pin_bytes < chunk_bytes + 16) {
            for (int i = 0; i < kPin; ++i) {
                if (pin[i]) cudaFreeHost(pin[i]);
                pin[i] = nullptr;
            }
            c->ply_pin_bytes = 0;
            for (int i = 0; i < kPin; ++i) CU(cudaMallocHost(&pin[i], chunk_bytes + 16));
            c->ply_pin_bytes = chunk_bytes + 16;
        }
        const uint64_t use_bytes = std::min(chunk_bytes, n * info.record_bytes);
        for (auto& d : dev) d = (uint8_t*)be.dmalloc(use_bytes + 16);
        for (auto& e : copied) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const int nthreads = (int)std::max<uint64_t>(1, std::min<uint64_t>({32, std::thread::hardware_concurrency(), nchunks * ppc}));
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(reader);
        for (uint64_t k = 0; k < nchunks; ++k) {
            auto t0 = std::chrono::steady_clock::now();
            {
                std::unique_lock<std::mutex> l(mu);
                cv_done.wait(l, [&] { return err || done[k] == chunk_pieces(k); });
                if (err) throw BuildError(PCV_ERR_IO, err == 1 ? "truncated PLY body" : "read failed");
            }
            ms_wait_read += since(t0);
            const uint64_t first = k * chunk_points, m = std::min(chunk_points, n - first);
            const int ps = (int)(k % kPin), ds = (int)(k % kDev);
            CU(cudaMemcpyAsync(dev[ds], pin[ps], m * info.record_bytes, cudaMemcpyHostToDevice, c->stream));
            CU(cudaEventRecord(copied[ps], c->stream));
            ply_launch(c, a, dev[ds], m, first, d_partial + (first / T) * 6);  // same stream: ordered after the copy
            if (k >= 1) {  // chunk k - 1 has left pinned memory (its copy precedes this one): its slot may take chunk k + 2
                t0 = std::chrono::steady_clock::now();
                CU(cudaEventSynchronize(copied[(k - 1) % kPin]));
                ms_wait_copy += since(t0);
                {
                    std::lock_guard<std::mutex> l(mu);
                    released = k - 1 + kPin + 1;
                }
                cv_slot.notify_all();
            }
        }
        auto t0 = std::chrono::steady_clock::now();
        ply_reduce_bbox(c, d_partial, ntiles, bmin, bmax);  // synchronises the stream
        if (timing)
            fprintf(stderr, "[ply_load] n=%llu chunks=%llu threads=%d wait-for-read %.1f ms wait-for-copy %.1f ms drain %.1f ms total %.1f ms (%.1f GB/s)\n",
                    (unsigned long long)n, (unsigned long long)nchunks, nthreads, ms_wait_read, ms_wait_copy, since(t0), since(t_begin),
                    n * info.record_bytes / since(t_begin) / 1e6);
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

int pcv_ply_read_header(const char* path, pcv_ply_info* out) {
    if (!path || !out) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    *out = ply_parse(path);
    return PCV_OK;
    API_CATCH
}

int pcv_ply_unpack_device(pcv_ctx* c, const pcv_ply_info* info, const void* dev_records, uint64_t n, double* dx, double* dy, double* dz,
                          uint8_t* drgb, float* dintensity, double bbox_min[3], double bbox_max[3]) {
    if (!c || !info || (n && (!dev_records || !dx || !dy || !dz))) return fail(PCV_ERR_INVALID, "null argument");
    if (reinterpret_cast<uintptr_t>(dev_records) & 15) return fail(PCV_ERR_INVALID, "records must be 16-byte aligned");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ply_validate(*info);
    PlyUnpackArgs a = ply_args(*info, dx, dy, dz, drgb, dintensity);
    const uint64_t ntiles = (n + a.tile_points - 1) / a.tile_points;
    double* d_partial = ntiles ? (double*)c->be->dmalloc(ntiles * 6 * 8) : nullptr;
    try {
        ply_launch(c, a, (const uint8_t*)dev_records, n, 0, d_partial);
        double mn[3], mx[3];
        ply_reduce_bbox(c, d_partial, ntiles, mn, mx);
        for (int k = 0; k < 3; ++k) {
            if (bbox_min) bbox_min[k] = mn[k];
            if (bbox_max) bbox_max[k] = mx[k];
        }
    } catch (...) {
        c->be->dfree(d_partial);
        throw;
    }
    c->be->dfree(d_partial);
    return PCV_OK;
    API_CATCH
}

int pcv_ply_load_device(pcv_ctx* c, const char* path, const pcv_ply_info* info, double* dx, double* dy, double* dz, uint8_t* drgb,
                        float* dintensity, double bbox_min[3], double bbox_max[3]) {
    if (!c || !path || !info || !bbox_min || !bbox_max) return fail(PCV_ERR_INVALID, "null argument");
    if (info->num_points && (!dx || !dy || !dz)) return fail(PCV_ERR_INVALID, "null output array");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    ply_load(c, path, *info, dx, dy, dz, drgb, dintensity, bbox_min, bbox_max);
    return PCV_OK;
    API_CATCH
}

int pcv_build_octree_from_file(pcv_ctx* c, const char* path, double resolution, int with_intensity, pcv_octree** out) {
    if (!c || !path || !out) return fail(PCV_ERR_INVALID, "null argument");
    *out = nullptr;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    const pcv_ply_info info = ply_parse(path);
    const uint64_t n = info.num_points;
    if (n >= 0xFFFFFFFFull) throw BuildError(PCV_ERR_UNSUPPORTED, "more than 2^32-1 points per context");
    if (n && !info.has_color) throw BuildError(PCV_ERR_INVALID, "color is mandatory (point counts come from .rgb, on_disk.rs:23-33)");
    if (n && with_intensity && !info.has_intensity) throw BuildError(PCV_ERR_INVALID, "the file has no float 'intensity' property");
    CudaBackend& be = *c->be;
    std::vector<void*> owned;
    auto alloc = [&](size_t bytes) {
        void* p = be.dmalloc(bytes + 16);
        owned.push_back(p);
        return p;
    };
    int rc;
    try {
        PointsView v{};
        v.n = n;
        v.stride = 1;
        double bmin[3], bmax[3];
        if (n) {
            double* xyz = (double*)alloc(n * 24);
            v.x = xyz, v.y = xyz + n, v.z = xyz + 2 * n;
            v.rgb = (uint8_t*)alloc(n * 3);
            v.intensity = with_intensity ? (float*)alloc(n * 4) : nullptr;
        }
        ply_load(c, path, info, const_cast<double*>(v.x), const_cast<double*>(v.y), const_cast<double*>(v.z), const_cast<uint8_t*>(v.rgb),
                 const_cast<float*>(v.intensity), bmin, bmax);
        rc = build_impl(c, v, resolution, bmin, bmax, out);
    } catch (...) {
        for (void* p : owned) be.dfree(p);
        throw;
    }
    for (void* p : owned) be.dfree(p);
    return rc;
    API_CATCH
}
