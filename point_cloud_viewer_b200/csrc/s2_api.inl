// s2_api.inl — the C ABI of the S2-cell point cloud (included by pcv_api.cu).
//
// What the reference does with files, per point and per batch (src/read_write/s2.rs:52-125: radius check, bounding box,
// CellID::from_point(p).parent(split_level), one NodeWriter per cell, S2Meta from get_meta), becomes: keys -> one stable sort by
// key -> cell-contiguous arrays in HBM.  Cells are kept in id order; inside a cell the points keep their input order, which
// is the order the reference's per-cell files end up in.  Queries: AllPoints and S2Cells(CellUnion) - the two
// PointLocation kinds whose cell selection (cells_intersecting_region with CellUnion::intersects_cell, src/s2_cells/mod.rs:157-168,
// 233-241) and point test (contains_cellid(from_point(p)), src/geometry/s2_cell_union.rs:27-31) are pure integer arithmetic.
// The polyhedron locations go through CellUnion::rect_bound / Rect::intersects_cell of the un-vendored s2 crate (latitude /
// longitude intervals: asin, atan2) and are not built.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "s2_disk.hpp"

struct pcv_s2cloud {
    pcv_ctx* ctx = nullptr;
    uint32_t level = 20;
    uint64_t n = 0;
    bool has_rgb = false, has_intensity = false;
    double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    std::vector<uint64_t> ids, counts, starts;  // cells sorted by id; starts = exclusive prefix of counts
    double* d_xyz = nullptr;                    // n * 3 (Encoding::Plain)
    uint8_t* d_rgb = nullptr;
    float* d_intensity = nullptr;
    uint32_t* d_src = nullptr;                  // input index of every slot
    float ms_device = 0.f;                      // CUDA events around the build's kernels (keys ... gather)
    uint32_t kernel_launches = 0;
};

namespace {

using namespace pcv;

struct RunStart {
    uint64_t key, start;
};
__global__ void __launch_bounds__(256) k_s2_run_starts(const uint64_t* __restrict__ keys, uint64_t n, RunStart* __restrict__ out, unsigned long long* __restrict__ cursor) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (i == 0 || keys[i] != keys[i - 1]) out[atomicAdd(cursor, 1ull)] = RunStart{keys[i], i};
}

inline uint32_t s2_grid(pcv_ctx* c, uint64_t n) { return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 16)); }

// S2Splitter::write over the whole cloud + get_meta.  `v` holds device pointers.
pcv_s2cloud* s2_build_core(pcv_ctx* c, const PointsView& v, uint32_t level) {
    std::unique_ptr<pcv_s2cloud> s(new pcv_s2cloud());
    s->ctx = c;
    s->level = level;
    s->n = v.n;
    s->has_rgb = v.rgb != nullptr;
    s->has_intensity = v.intensity != nullptr;
    const uint64_t n = v.n;
    if (n == 0) return s.release();
    Scratch sc(c);
    uint64_t* keys = sc.alloc<uint64_t>(n);
    uint64_t* keys2 = sc.alloc<uint64_t>(n);
    uint32_t* idx = sc.alloc<uint32_t>(n);
    uint32_t* order = (uint32_t*)c->be->dmalloc(n * 4);  // becomes d_src
    struct Own {
        pcv_ctx* c;
        std::vector<void*> p;
        ~Own() {
            for (void* q : p) c->be->dfree(q);
        }
        void release() { p.clear(); }
    } own{c, {order}};
    unsigned long long* scal = sc.alloc<unsigned long long>(3);  // first bad index, run count, run cursor
    const unsigned long long init[3] = {~0ull, 0ull, 0ull};
    c->be->h2d(scal, init, sizeof init);
    struct Ev {
        cudaEvent_t e = nullptr;
        ~Ev() {
            if (e) cudaEventDestroy(e);
        }
    } ev0, ev1;
    CU(cudaEventCreate(&ev0.e));
    CU(cudaEventCreate(&ev1.e));
    const uint64_t l0 = c->be->launches;
    CU(cudaEventRecord(ev0.e, c->stream));
    k_s2_keys<<<s2_grid(c, n), 256, 0, c->stream>>>(v, (int)level, keys, idx, scal);
    c->be->launches++;
    CU(cudaGetLastError());
    unsigned long long bad = 0;
    c->be->d2h(&bad, scal, 8);
    if (bad != ~0ull) {
        double p[3];
        c->be->d2h(&p[0], v.x + bad * v.stride, 8);
        c->be->d2h(&p[1], v.y + bad * v.stride, 8);
        c->be->d2h(&p[2], v.z + bad * v.stride, 8);
        char msg[200];
        snprintf(msg, sizeof msg, "Point (%g, %g, %g) is not a valid ECEF point", p[0], p[1], p[2]);  // read_write/s2.rs:66-70
        throw BuildError(PCV_ERR_INVALID, msg);
    }
    c->be->bbox(v, s->bmin, s->bmax);
    // stable sort of (key, index) by key: the bits below the level's lsb are zero and the lsb itself is set in every key
    const int begin_bit = 2 * (kS2MaxLevel - (int)level) + 1;
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, order, (int64_t)n, begin_bit, 64, c->stream));
    void* tmp = sc.alloc<uint8_t>(tmp_bytes ? tmp_bytes : 16);
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, order, (int64_t)n, begin_bit, 64, c->stream));
    c->be->launches += 8;  // the library's histogram + onesweep passes (an estimate: they are not this repository's kernels)
    // cells: every run start of the sorted keys
    k_s2_count_runs<<<s2_grid(c, n), 256, 0, c->stream>>>(keys2, n, scal + 1);
    c->be->launches++;
    unsigned long long nruns = 0;
    c->be->d2h(&nruns, scal + 1, 8);
    RunStart* runs = sc.alloc<RunStart>(nruns);
    k_s2_run_starts<<<s2_grid(c, n), 256, 0, c->stream>>>(keys2, n, runs, scal + 2);
    c->be->launches++;
    CU(cudaGetLastError());
    std::vector<RunStart> hr(nruns);
    c->be->d2h(hr.data(), runs, nruns * sizeof(RunStart));
    std::sort(hr.begin(), hr.end(), [](const RunStart& a, const RunStart& b) { return a.start < b.start; });
    s->ids.resize(nruns);
    s->counts.resize(nruns);
    s->starts.resize(nruns);
    for (size_t k = 0; k < nruns; ++k) {
        s->ids[k] = hr[k].key;
        s->starts[k] = hr[k].start;
        s->counts[k] = (k + 1 < nruns ? hr[k + 1].start : n) - hr[k].start;
        if (k && !(hr[k - 1].key < hr[k].key)) throw BuildError(PCV_ERR_CUDA, "S2 split: cell keys are not sorted");
    }
    // gather into cell-contiguous arrays
    S2GatherArgs g{};
    g.p = v;
    g.order = order;
    g.xyz = (double*)c->be->dmalloc(n * 24);
    own.p.push_back(g.xyz);
    if (v.rgb) {
        g.rgb = (uint8_t*)c->be->dmalloc(n * 3);
        own.p.push_back(g.rgb);
    }
    if (v.intensity) {
        g.intensity = (float*)c->be->dmalloc(n * 4);
        own.p.push_back(g.intensity);
    }
    k_s2_gather<<<s2_grid(c, n), 256, 0, c->stream>>>(g);
    c->be->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ev1.e, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&s->ms_device, ev0.e, ev1.e);
    s->kernel_launches = (uint32_t)(c->be->launches - l0);
    s->d_xyz = g.xyz;
    s->d_rgb = g.rgb;
    s->d_intensity = g.intensity;
    s->d_src = order;
    own.release();
    return s.release();
}

int s2_find_cell(const pcv_s2cloud* s, uint64_t id) {
    const auto it = std::lower_bound(s->ids.begin(), s->ids.end(), id);
    return it != s->ids.end() && *it == id ? (int)(it - s->ids.begin()) : -1;
}

// the union a caller passes, normalised (CellUnion::normalize) and validated
std::vector<uint64_t> s2_union_of(const uint64_t* ids, uint32_t n) {
    std::vector<uint64_t> u(ids, ids + n);
    for (uint64_t id : u)
        if (!s2_is_valid(id)) throw BuildError(PCV_ERR_INVALID, "invalid S2 cell id in the union");
    s2_normalize(u);
    return u;
}

}  // namespace

extern "C" {

int pcv_s2_cell_ids(pcv_ctx* c, const pcv_points* hp, uint32_t level, uint64_t* ids_out) {
    if (!c || !hp || (hp->n && !ids_out) || level > 30) return fail(PCV_ERR_INVALID, "null argument or level > 30");
    if (hp->n == 0) return PCV_OK;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    std::vector<void*> owned;
    pcv_points tmp = *hp;
    tmp.rgb = nullptr;
    tmp.intensity = nullptr;
    struct Free {
        pcv_ctx* c;
        std::vector<void*>& o;
        ~Free() {
            for (void* p : o) c->be->dfree(p);
        }
    } fr{c, owned};
    const PointsView v = stage_points(c, &tmp, owned);
    Scratch sc(c);
    uint64_t* keys = sc.alloc<uint64_t>(v.n);
    k_s2_keys<<<s2_grid(c, v.n), 256, 0, c->stream>>>(v, (int)level, keys, nullptr, nullptr);
    c->be->launches++;
    CU(cudaGetLastError());
    c->be->d2h(ids_out, keys, v.n * 8);
    return PCV_OK;
    API_CATCH
}

int pcv_s2_build_device(pcv_ctx* c, const pcv_points* dp, uint32_t split_level, pcv_s2cloud** out) {
    if (!c || !dp || !out || split_level > 30) return fail(PCV_ERR_INVALID, "null argument or split level > 30");
    if (dp->n > 0xFFFFFFFEull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-2 points per context");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    *out = s2_build_core(c, view_of(dp), split_level);
    return PCV_OK;
    API_CATCH
}

int pcv_s2_build(pcv_ctx* c, const pcv_points* hp, uint32_t split_level, pcv_s2cloud** out) {
    if (!c || !hp || !out || split_level > 30) return fail(PCV_ERR_INVALID, "null argument or split level > 30");
    if (hp->n > 0xFFFFFFFEull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-2 points per context");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    std::vector<void*> owned;
    struct Free {
        pcv_ctx* c;
        std::vector<void*>& o;
        ~Free() {
            for (void* p : o) c->be->dfree(p);
        }
    } fr{c, owned};
    const PointsView v = stage_points(c, hp, owned);
    *out = s2_build_core(c, v, split_level);
    return PCV_OK;
    API_CATCH
}

void pcv_s2_free(pcv_s2cloud* s) {
    if (!s) return;
    pcv_ctx* c = s->ctx;
    {
        std::lock_guard<std::mutex> g(c->mu);
        cudaSetDevice(c->device);
        c->be->dfree(s->d_xyz);
        c->be->dfree(s->d_rgb);
        c->be->dfree(s->d_intensity);
        c->be->dfree(s->d_src);
    }
    delete s;
}

int pcv_s2_info(const pcv_s2cloud* s, uint64_t* num_cells, uint64_t* num_points, uint32_t* split_level, double bbox_min[3], double bbox_max[3], int* has_color,
                int* has_intensity) {
    if (!s) return fail(PCV_ERR_INVALID, "null argument");
    if (num_cells) *num_cells = s->ids.size();
    if (num_points) *num_points = s->n;
    if (split_level) *split_level = s->level;
    for (int a = 0; a < 3; ++a) {
        if (bbox_min) bbox_min[a] = s->bmin[a];
        if (bbox_max) bbox_max[a] = s->bmax[a];
    }
    if (has_color) *has_color = s->has_rgb ? 1 : 0;
    if (has_intensity) *has_intensity = s->has_intensity ? 1 : 0;
    return PCV_OK;
}

int pcv_s2_build_stats(const pcv_s2cloud* s, float* ms_device, uint32_t* kernel_launches, uint64_t* algorithmic_bytes) {
    if (!s) return fail(PCV_ERR_INVALID, "null argument");
    if (ms_device) *ms_device = s->ms_device;
    if (kernel_launches) *kernel_launches = s->kernel_launches;
    // every input point read once (24 + 3 + 4 B) and written once into its cell (the same bytes)
    if (algorithmic_bytes) *algorithmic_bytes = 2ull * s->n * (24ull + (s->has_rgb ? 3ull : 0ull) + (s->has_intensity ? 4ull : 0ull));
    return PCV_OK;
}

int pcv_s2_cells(const pcv_s2cloud* s, uint64_t* ids_out, uint64_t* num_points_out) {
    if (!s) return fail(PCV_ERR_INVALID, "null argument");
    for (size_t k = 0; k < s->ids.size(); ++k) {
        if (ids_out) ids_out[k] = s->ids[k];
        if (num_points_out) num_points_out[k] = s->counts[k];
    }
    return PCV_OK;
}

int pcv_s2_cell_data(const pcv_s2cloud* s, uint64_t cell_id, double* xyz_out, uint8_t* rgb_out, float* intensity_out, uint64_t* src_index_out) {
    if (!s) return fail(PCV_ERR_INVALID, "null argument");
    const int k = s2_find_cell(s, cell_id);
    if (k < 0) return fail(PCV_ERR_NOT_FOUND, "Could not get cell %s.", s2_to_token(cell_id).c_str());
    if ((rgb_out && !s->d_rgb) || (intensity_out && !s->d_intensity)) return fail(PCV_ERR_INVALID, "the cloud does not hold that attribute");
    API_TRY
    pcv_ctx* c = s->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    const uint64_t st = s->starts[k], n = s->counts[k];
    if (xyz_out) c->be->d2h(xyz_out, s->d_xyz + 3 * st, n * 24);
    if (rgb_out) c->be->d2h(rgb_out, s->d_rgb + 3 * st, n * 3);
    if (intensity_out) c->be->d2h(intensity_out, s->d_intensity + st, n * 4);
    if (src_index_out) {
        std::vector<uint32_t> t(n);
        c->be->d2h(t.data(), s->d_src + st, n * 4);
        for (uint64_t i = 0; i < n; ++i) src_index_out[i] = t[i];
    }
    return PCV_OK;
    API_CATCH
}

int pcv_s2_cells_in_union(const pcv_s2cloud* s, const uint64_t* union_ids, uint32_t n_union, uint64_t* ids_out, uint64_t cap, uint64_t* n_out) {
    if (!s || !n_out || (n_union && !union_ids)) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    uint64_t n = 0;
    if (!union_ids) {  // PointLocation::AllPoints: every cell (mod.rs:160)
        for (uint64_t id : s->ids) {
            if (ids_out && n < cap) ids_out[n] = id;
            ++n;
        }
    } else {
        const std::vector<uint64_t> u = s2_union_of(union_ids, n_union);
        for (uint64_t id : s->ids)
            if (s2_union_intersects(u.data(), (uint32_t)u.size(), id)) {  // Region::intersects_cell for a CellUnion = intersects_cellid(cell.id)
                if (ids_out && n < cap) ids_out[n] = id;
                ++n;
            }
    }
    *n_out = n;
    return PCV_OK;
    API_CATCH
}

int pcv_s2_query_union(const pcv_s2cloud* s, const uint64_t* union_ids, uint32_t n_union, double* xyz_out, uint8_t* rgb_out, float* intensity_out,
                       uint64_t* src_index_out, uint64_t cap, uint64_t* n_out, uint64_t* tested_out) {
    if (!s || !n_out || (n_union && !union_ids)) return fail(PCV_ERR_INVALID, "null argument");
    if ((rgb_out && !s->d_rgb) || (intensity_out && !s->d_intensity)) return fail(PCV_ERR_INVALID, "the cloud does not hold that attribute");
    API_TRY
    pcv_ctx* c = s->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    *n_out = 0;
    if (tested_out) *tested_out = 0;
    if (s->n == 0) return PCV_OK;
    // the cells the location selects form one slot range [lo, hi) per contiguous run of ids; take the hull of all of them:
    // a point outside the selected cells cannot be in the union, so testing it changes nothing but the tested count
    uint64_t lo = s->n, hi = 0, tested = 0;
    std::vector<uint64_t> u;
    if (union_ids) u = s2_union_of(union_ids, n_union);
    std::vector<std::pair<uint64_t, uint64_t>> segs;  // selected cells as slot ranges, merged
    for (size_t k = 0; k < s->ids.size(); ++k)
        if (!union_ids || s2_union_intersects(u.data(), (uint32_t)u.size(), s->ids[k])) {
            const uint64_t a = s->starts[k], b = a + s->counts[k];
            if (!segs.empty() && segs.back().second == a)
                segs.back().second = b;
            else
                segs.emplace_back(a, b);
            tested += s->counts[k];
            lo = std::min(lo, a);
            hi = std::max(hi, b);
        }
    if (tested_out) *tested_out = tested;
    if (segs.empty()) return PCV_OK;
    Scratch sc(c);
    uint64_t total = 0;
    uint64_t* slots = nullptr;
    if (!union_ids) {  // AllPoints: every slot survives, no test
        total = s->n;
    } else {
        const uint64_t* du = sc.upload(u.data(), u.size());
        PointsView pv{};
        pv.x = s->d_xyz;
        pv.y = s->d_xyz + 1;
        pv.z = s->d_xyz + 2;
        pv.stride = 3;
        pv.n = s->n;
        uint8_t* flag = sc.alloc<uint8_t>(hi - lo);
        CU(cudaMemsetAsync(flag, 0, hi - lo, c->stream));
        for (const auto& sg : segs) {
            k_s2_union_mask<<<s2_grid(c, sg.second - sg.first), 256, 0, c->stream>>>(pv, sg.first, sg.second - sg.first, du, (uint32_t)u.size(), flag + (sg.first - lo));
            c->be->launches++;
        }
        CU(cudaGetLastError());
        slots = sc.alloc<uint64_t>(hi - lo);
        long long* dsel = sc.alloc<long long>(1);
        thrust::counting_iterator<uint64_t> first(lo);
        size_t tb = 0;
        CU(cub::DeviceSelect::Flagged(nullptr, tb, first, flag, slots, dsel, (int64_t)(hi - lo), c->stream));
        void* tmp = sc.alloc<uint8_t>(tb ? tb : 16);
        CU(cub::DeviceSelect::Flagged(tmp, tb, first, flag, slots, dsel, (int64_t)(hi - lo), c->stream));
        c->be->launches += 2;
        long long sel = 0;
        c->be->d2h(&sel, dsel, 8);
        total = (uint64_t)sel;
    }
    *n_out = total;
    const uint64_t emit = std::min<uint64_t>(total, cap);
    if (emit == 0) return PCV_OK;
    if (!union_ids) {  // straight copies
        if (xyz_out) c->be->d2h(xyz_out, s->d_xyz, emit * 24);
        if (rgb_out) c->be->d2h(rgb_out, s->d_rgb, emit * 3);
        if (intensity_out) c->be->d2h(intensity_out, s->d_intensity, emit * 4);
        if (src_index_out) {
            std::vector<uint32_t> t(emit);
            c->be->d2h(t.data(), s->d_src, emit * 4);
            for (uint64_t i = 0; i < emit; ++i) src_index_out[i] = t[i];
        }
        return PCV_OK;
    }
    S2EmitArgs e{};
    e.slots = slots;
    e.n = emit;
    e.xyz = s->d_xyz;
    e.rgb = s->d_rgb;
    e.intensity = s->d_intensity;
    e.src = s->d_src;
    e.xyz_out = sc.alloc<double>(emit * 3);
    e.rgb_out = rgb_out ? sc.alloc<uint8_t>(emit * 3) : nullptr;
    e.intensity_out = intensity_out ? sc.alloc<float>(emit) : nullptr;
    e.src_out = src_index_out ? sc.alloc<uint64_t>(emit) : nullptr;
    k_s2_emit<<<s2_grid(c, emit), 256, 0, c->stream>>>(e);
    c->be->launches++;
    CU(cudaGetLastError());
    if (xyz_out) c->be->d2h(xyz_out, e.xyz_out, emit * 24);
    if (rgb_out) c->be->d2h(rgb_out, e.rgb_out, emit * 3);
    if (intensity_out) c->be->d2h(intensity_out, e.intensity_out, emit * 4);
    if (src_index_out) c->be->d2h(src_index_out, e.src_out, emit * 8);
    return PCV_OK;
    API_CATCH
}

// What S2Splitter<RawNodeWriter> + S2Meta::to_proto leave in a directory (s2_disk.hpp).
int pcv_s2_write_dir(const pcv_s2cloud* s, const char* dir) {
    if (!s || !dir) return fail(PCV_ERR_INVALID, "null argument");
    if (s->n == 0) return fail(PCV_ERR_INVALID, "an S2 cloud without points has no meta (S2Splitter::get_meta returns None)");
    API_TRY
    pcv_ctx* c = s->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    mkdir(dir, 0777);
    const std::string base = std::string(dir) + "/";
    std::vector<uint8_t> buf;
    for (size_t k = 0; k < s->ids.size(); ++k) {
        const uint64_t st = s->starts[k], n = s->counts[k];
        const std::string stem = base + s2_to_token(s->ids[k]);
        buf.resize(n * 24);
        c->be->d2h(buf.data(), s->d_xyz + 3 * st, n * 24);
        if (!write_whole_file(stem + ".xyz", buf.data(), n * 24)) return fail(PCV_ERR_IO, "cannot write %s.xyz", stem.c_str());
        if (s->d_rgb) {
            c->be->d2h(buf.data(), s->d_rgb + 3 * st, n * 3);
            if (!write_whole_file(stem + ".rgb", buf.data(), n * 3)) return fail(PCV_ERR_IO, "cannot write %s.rgb", stem.c_str());
        }
        if (s->d_intensity) {
            c->be->d2h(buf.data(), s->d_intensity + st, n * 4);
            if (!write_whole_file(stem + ".intensity", buf.data(), n * 4)) return fail(PCV_ERR_IO, "cannot write %s.intensity", stem.c_str());
        }
    }
    S2MetaData m;
    for (int a = 0; a < 3; ++a) m.bbox_min[a] = s->bmin[a], m.bbox_max[a] = s->bmax[a];
    m.ids = s->ids;
    m.counts = s->counts;
    m.has_color = s->has_rgb;
    m.has_intensity = s->has_intensity;
    const std::string meta = encode_s2_meta(m);
    if (!write_whole_file(base + "meta.pb", meta.data(), meta.size())) return fail(PCV_ERR_IO, "cannot write %smeta.pb", base.c_str());
    return PCV_OK;
    API_CATCH
}

// S2Cells::from_data_provider over an OnDiskDataProvider (s2_cells/mod.rs:203-216, data_provider/on_disk.rs).
int pcv_s2_load_dir(pcv_ctx* c, const char* dir, pcv_s2cloud** out) {
    if (!c || !dir || !out) return fail(PCV_ERR_INVALID, "null argument");
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    const std::string base = std::string(dir) + "/";
    std::string raw;
    if (!read_whole_file(base + "meta.pb", raw)) return fail(PCV_ERR_IO, "cannot read %smeta.pb", base.c_str());
    S2MetaData m;
    int version = 0;
    const std::string err = decode_s2_meta(raw, m, version);
    if (!err.empty()) return fail(PCV_ERR_INVALID, "%s", err.c_str());
    // cells in id order (the proto carries them in hash-map order)
    std::vector<size_t> order(m.ids.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return m.ids[a] < m.ids[b]; });
    std::unique_ptr<pcv_s2cloud> s(new pcv_s2cloud());
    s->ctx = c;
    s->has_rgb = m.has_color;
    s->has_intensity = m.has_intensity;
    for (int a = 0; a < 3; ++a) s->bmin[a] = m.bbox_min[a], s->bmax[a] = m.bbox_max[a];
    uint64_t n = 0;
    int level = -1;
    for (size_t k : order) {
        if (!s2_is_valid(m.ids[k])) return fail(PCV_ERR_INVALID, "invalid S2 cell id %llx in meta.pb", (unsigned long long)m.ids[k]);
        if (!s->ids.empty() && s->ids.back() == m.ids[k]) return fail(PCV_ERR_INVALID, "cell %s is listed twice", s2_to_token(m.ids[k]).c_str());
        s->ids.push_back(m.ids[k]);
        s->counts.push_back(m.counts[k]);
        s->starts.push_back(n);
        n += m.counts[k];
        level = std::max(level, s2_level(m.ids[k]));
    }
    if (n > 0xFFFFFFFEull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-2 points per context");
    s->n = n;
    s->level = (uint32_t)std::max(level, 0);
    if (n == 0) {
        *out = s.release();
        return PCV_OK;
    }
    struct Own {
        pcv_ctx* c;
        std::vector<void*> p;
        ~Own() {
            for (void* q : p) c->be->dfree(q);
        }
    } own{c, {}};
    double* dx = (double*)c->be->dmalloc(n * 24);
    own.p.push_back(dx);
    uint8_t* dr = nullptr;
    float* di = nullptr;
    uint32_t* ds = (uint32_t*)c->be->dmalloc(n * 4);
    own.p.push_back(ds);
    if (m.has_color) {
        dr = (uint8_t*)c->be->dmalloc(n * 3);
        own.p.push_back(dr);
    }
    if (m.has_intensity) {
        di = (float*)c->be->dmalloc(n * 4);
        own.p.push_back(di);
    }
    std::string data;
    std::vector<uint32_t> iota;
    for (size_t k = 0; k < s->ids.size(); ++k) {
        const uint64_t st = s->starts[k], cnt = s->counts[k];
        const std::string stem = base + s2_to_token(s->ids[k]);
        auto load = [&](const char* ext, void* dst, uint64_t bytes) -> bool {
            if (!read_whole_file(stem + ext, data) || data.size() != bytes) return false;
            if (bytes) c->be->h2d(dst, data.data(), bytes);
            return true;
        };
        if (!load(".xyz", dx + 3 * st, cnt * 24)) return fail(PCV_ERR_NOT_FOUND, "Could not read %llu points of cell %s.xyz", (unsigned long long)cnt, stem.c_str());
        if (dr && !load(".rgb", dr + 3 * st, cnt * 3)) return fail(PCV_ERR_NOT_FOUND, "Could not read cell %s.rgb", stem.c_str());
        if (di && !load(".intensity", di + st, cnt * 4)) return fail(PCV_ERR_NOT_FOUND, "Could not read cell %s.intensity", stem.c_str());
        iota.resize(cnt);
        for (uint64_t i = 0; i < cnt; ++i) iota[i] = (uint32_t)(st + i);  // no provenance on disk: the slot number
        if (cnt) c->be->h2d(ds + st, iota.data(), cnt * 4);
    }
    CU(cudaStreamSynchronize(c->stream));
    s->d_xyz = dx;
    s->d_rgb = dr;
    s->d_intensity = di;
    s->d_src = ds;
    own.p.clear();
    *out = s.release();
    return PCV_OK;
    API_CATCH
}

int pcv_s2_union_contains(pcv_ctx* c, const pcv_points* hp, const uint64_t* union_ids, uint32_t n_union, uint8_t* mask_out) {
    if (!c || !hp || (hp->n && !mask_out) || (n_union && !union_ids)) return fail(PCV_ERR_INVALID, "null argument");
    if (hp->n == 0) return PCV_OK;
    API_TRY
    std::lock_guard<std::mutex> g(c->mu);
    CU(cudaSetDevice(c->device));
    const std::vector<uint64_t> u = s2_union_of(union_ids, n_union);
    std::vector<void*> owned;
    struct Free {
        pcv_ctx* c;
        std::vector<void*>& o;
        ~Free() {
            for (void* p : o) c->be->dfree(p);
        }
    } fr{c, owned};
    pcv_points tmp = *hp;
    tmp.rgb = nullptr;
    tmp.intensity = nullptr;
    const PointsView v = stage_points(c, &tmp, owned);
    Scratch sc(c);
    const uint64_t* du = sc.upload(u.data(), u.size());
    uint8_t* flag = sc.alloc<uint8_t>(v.n);
    k_s2_union_mask<<<s2_grid(c, v.n), 256, 0, c->stream>>>(v, 0, v.n, du, (uint32_t)u.size(), flag);
    c->be->launches++;
    CU(cudaGetLastError());
    c->be->d2h(mask_out, flag, v.n);
    return PCV_OK;
    API_CATCH
}

}  // extern "C"
