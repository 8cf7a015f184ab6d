// kernels_build.cuh — sm_100a kernels of the octree build and the CUDA Backend that drives them.
//
//   k_bbox      a1  find_bounding_box (generation.rs:256-270): streaming min/max, warp-shuffle reduce
//   k_hist      a3-a6  per tile: decode -> G descent steps (chain.h) -> digit histogram in shared memory
//   k_scan_*    per-digit exclusive prefix over the tiles of each active node (+ node totals)
//   k_scatter   a4  stable multi-way partition of a tile (warp match ranking), re-running the descent
//   k_place     a7  closed-form LOD subsampling + up-chain re-encode + final node-contiguous store
//
// All of them stream SoA/record arrays once with coalesced accesses; none has a dense contraction,
// so there is no tensor-core path here.  The descent is FP64 (IEEE divide + FMA) by definition of
// the reference's codec, compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "build_host.hpp"
#include "chain_device.cuh"

namespace pcv {

#define PCV_CUDA_CHECK(x)                                                                            \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) throw BuildError(-2, std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// bbox
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// partial[block][6] = min xyz, max xyz.  Grid-stride, 4 independent loads in flight per coordinate.
__global__ void __launch_bounds__(256) k_bbox(PointsView p, double* __restrict__ partial) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t s = p.stride;
    for (; i + 3 * step < p.n; i += 4 * step) {
        double vx[4], vy[4], vz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            vx[u] = __ldg(p.x + (i + u * step) * s);
            vy[u] = __ldg(p.y + (i + u * step) * s);
            vz[u] = __ldg(p.z + (i + u * step) * s);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            mn[0] = fmin(mn[0], vx[u]);
            mx[0] = fmax(mx[0], vx[u]);
            mn[1] = fmin(mn[1], vy[u]);
            mx[1] = fmax(mx[1], vy[u]);
            mn[2] = fmin(mn[2], vz[u]);
            mx[2] = fmax(mx[2], vz[u]);
        }
    }
    for (; i < p.n; i += step) {
        double x = __ldg(p.x + i * s), y = __ldg(p.y + i * s), z = __ldg(p.z + i * s);
        mn[0] = fmin(mn[0], x);
        mx[0] = fmax(mx[0], x);
        mn[1] = fmin(mn[1], y);
        mx[1] = fmax(mx[1], y);
        mn[2] = fmin(mn[2], z);
        mx[2] = fmax(mx[2], z);
    }
    __shared__ double sh[8][6];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double lo = warp_min(mn[a]), hi = warp_max(mx[a]);
        if (l == 0) {
            sh[w][a] = lo;
            sh[w][3 + a] = hi;
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sh[0][threadIdx.x];
        for (int k = 1; k < 8; ++k) v = threadIdx.x < 3 ? fmin(v, sh[k][threadIdx.x]) : fmax(v, sh[k][threadIdx.x]);
        partial[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// L2 prefetch of the tile the SM will work on one wave later
// ------------------------------------------------------------------------------------------------
// k_place streams leaf tiles whose addresses are known long before they are needed.  One thread asks the TMA unit to
// pull the tile that is about one wave of resident blocks away into L2 (cp.async.bulk.prefetch.L2: no registers, no L1
// lines, no completion to wait for), so that when that block runs its loads hit L2 instead of HBM (measured at N = 1e9:
// k_place 13.4 -> 11.2 ms; the same prefetch did not help k_hist / k_scatter, whose time is not load latency).  The byte
// range is shrunk to 16-byte alignment inside [p, p + bytes): nothing outside the caller's range is ever touched.
__device__ __forceinline__ void l2_prefetch(const void* p, uint64_t bytes) {
    const uintptr_t b = (reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15;
    const uintptr_t e = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(uintptr_t)15;
    if (e > b) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b), "r"((uint32_t)(e - b)) : "memory");
}
// ------------------------------------------------------------------------------------------------
// record load / store helpers
// ------------------------------------------------------------------------------------------------
template <bool WIDE>
struct RecT;
template <>
struct RecT<false> {
    typedef RecN type;
};
template <>
struct RecT<true> {
    typedef RecW type;
};

template <bool WIDE>
__device__ __forceinline__ void load_rec(const void* base, uint64_t i, uint64_t c[3], uint32_t& idx) {
    if (WIDE) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(base) + 2 * i;  // 32-byte record
        const ulonglong2 v0 = __ldg(p), v1 = __ldg(p + 1);
        c[0] = v0.x;
        c[1] = v0.y;
        c[2] = v1.x;
        idx = (uint32_t)v1.y;
    } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base) + i);  // 16-byte record
        c[0] = v.x;
        c[1] = v.y;
        c[2] = v.z;
        idx = v.w;
    }
}
template <bool WIDE>
__device__ __forceinline__ void store_rec(void* base, uint64_t i, const uint64_t c[3], uint32_t idx) {
    if (WIDE) {
        ulonglong2 v0, v1;
        v0.x = c[0];
        v0.y = c[1];
        v1.x = c[2];
        v1.y = idx;
        ulonglong2* p = reinterpret_cast<ulonglong2*>(base) + 2 * i;
        p[0] = v0;
        p[1] = v1;
    } else {
        uint4 v;
        v.x = (uint32_t)c[0];
        v.y = (uint32_t)c[1];
        v.z = (uint32_t)c[2];
        v.w = idx;
        reinterpret_cast<uint4*>(base)[i] = v;
    }
}

// Position of item `i` of a tile as the node's file would hand it to split(): raw for the root,
// decoded from the node's own encoding otherwise (raw.rs:127-216).
template <bool ROOT, bool WIDE>
__device__ __forceinline__ void load_position(const PassArgs& a, const TileDesc& t, const ActiveDesc& act, uint32_t i,
                                              double q[3], uint32_t& idx) {
    const uint64_t g = t.start + i;
    if (ROOT) {
        q[0] = __ldg(a.pts.x + g * a.pts.stride);
        q[1] = __ldg(a.pts.y + g * a.pts.stride);
        q[2] = __ldg(a.pts.z + g * a.pts.stride);
        idx = (uint32_t)g;
    } else {
        uint64_t c[3];
        load_rec<WIDE>(a.rec_in, g, c, idx);
        const int enc = a.lv.enc[a.level];
        if (a.lv.fast) {
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = decode1_fast(c[k], act.m[k], act.e, enc);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = decode1(c[k], act.m[k], act.e, enc);
        }
    }
}

// One descent step through the exact fast paths when the level table admits them (identical results).
__device__ __forceinline__ Step descend_any(const LevelTable& lv, int child_level, double q[3], double m[3], double e_cur) {
    if (lv.fast) return descend_fast(q, m, e_cur, lv.edge[child_level], lv.ry[child_level], lv.enc[child_level]);
    return descend(q, m, e_cur, lv.edge[child_level], lv.enc[child_level]);
}

template <bool ROOT>
__device__ __forceinline__ uint32_t load_colour(const PassArgs& a, uint64_t g) {
    if (ROOT) {
        const uint8_t* p = a.pts.rgb + 3 * g;
        return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
    }
    return __ldg(a.col_in + g);
}

// ------------------------------------------------------------------------------------------------
// hist
// ------------------------------------------------------------------------------------------------
// Position of item i as the node's file would hand it to split(): raw for the root, decoded from the node's own
// encoding (ENC_IN) otherwise.
template <bool ROOT, bool WIDE, int ENC_IN>
__device__ __forceinline__ void load_position_t(const PassArgs& a, const TileDesc& t, const ActiveDesc& act, uint32_t i, double q[3], uint32_t& idx) {
    const uint64_t g = t.start + i;
    if (ROOT) {
        // streamed once: bypass L1 (ld.global.cg).  The scatter kernel leaves only ~29 KB of L1 next to its shared
        // memory, too few lines for the loads in flight if they allocated there.
        q[0] = __ldcg(a.pts.x + g * a.pts.stride);
        q[1] = __ldcg(a.pts.y + g * a.pts.stride);
        q[2] = __ldcg(a.pts.z + g * a.pts.stride);
        idx = (uint32_t)g;
    } else {
        uint64_t c[3];
        load_rec<WIDE>(a.rec_in, g, c, idx);
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = decode_axis<ENC_IN>(c[k], act.m[k], act.e);
    }
}

// Points per thread in flight (x 3 independent axis chains each).  Measured at N = 1e9: one point per thread at 32
// registers (8 blocks per SM, 100 % occupancy) 17.2 ms; two points at 64 registers (4 blocks) 20.8 ms; two at 40 / 48
// registers (6 / 5 blocks, spilling) 18.3 / 19.1 ms - thread-level parallelism hides the load latency better than ILP.
constexpr int kHistItems = 1;

// Histogram of the G-level digits of one tile.  The last level needs only the child digit (no encode/decode).
template <bool ROOT, bool WIDE, int G, int FAST>
__device__ __forceinline__ unsigned hist_tile(const PassArgs& a, const TileDesc& t, const ActiveDesc& act, uint32_t* sh_hist) {
    unsigned bad = 0;
    for (uint32_t i0 = threadIdx.x; i0 < t.count; i0 += blockDim.x * kHistItems) {
        double q[kHistItems][3], m[kHistItems][3];
        uint32_t idx;
#pragma unroll
        for (int u = 0; u < kHistItems; ++u) {
            const uint32_t i = min(i0 + u * blockDim.x, t.count - 1);  // clamp: out-of-range lanes redo the last item, not counted
            if (ROOT) {
                load_position_t<true, WIDE, ENC_F64>(a, t, act, i, q[u], idx);
                if (FAST == 2) bad |= input_bad(q[u][0]) | input_bad(q[u][1]) | input_bad(q[u][2]);
            } else {
                PCV_ENC_SWITCH(a.lv.enc[a.level], load_position_t<false, WIDE, ENC>(a, t, act, i, q[u], idx);)
            }
            m[u][0] = act.m[0];
            m[u][1] = act.m[1];
            m[u][2] = act.m[2];
        }
        unsigned bin[kHistItems];
#pragma unroll
        for (int u = 0; u < kHistItems; ++u) bin[u] = 0;
        double e = act.e;
#pragma unroll
        for (int j = 1; j < G; ++j) {
            const double eh = a.lv.edge[a.level + j], ry = a.lv.ry[a.level + j];
            PCV_ENC_SWITCH(a.lv.enc[a.level + j], _Pragma("unroll") for (int u = 0; u < kHistItems; ++u) {
                uint32_t code[3];
                uint64_t codew[3];
                const unsigned d = WIDE ? level_step<ENC, FAST, true>(q[u], m[u], e, eh, ry, codew, bad) : level_step<ENC, FAST, true>(q[u], m[u], e, eh, ry, code, bad);
                bin[u] = (bin[u] << 3) | d;
            })
            e = eh;
        }
#pragma unroll
        for (int u = 0; u < kHistItems; ++u) {
            bin[u] = (bin[u] << 3) | level_digit(q[u], m[u], e);
            if (i0 + u * blockDim.x < t.count) atomicAdd(&sh_hist[bin[u]], 1u);
        }
    }
    return bad;
}

template <bool ROOT, bool WIDE, int G>
__global__ void __launch_bounds__(256, 8) k_hist(const __grid_constant__ PassArgs a) {
    extern __shared__ uint32_t sh_hist[];
    const TileDesc t = tile_of(a, blockIdx.x);
    const ActiveDesc act = a.d_active[t.active];
    constexpr int NB = 1 << (3 * G);
    for (int b = threadIdx.x; b < NB; b += blockDim.x) sh_hist[b] = 0;
    __syncthreads();
    if (a.lv.fast) {
        const unsigned bad = a.lv.fast == 2 ? hist_tile<ROOT, WIDE, G, 2>(a, t, act, sh_hist) : hist_tile<ROOT, WIDE, G, 1>(a, t, act, sh_hist);
        if (__syncthreads_or((int)bad)) {  // a numerator outside the proven range: redo the tile with the IEEE operator
            for (int b = threadIdx.x; b < NB; b += blockDim.x) sh_hist[b] = 0;
            __syncthreads();
            hist_tile<ROOT, WIDE, G, 0>(a, t, act, sh_hist);
        }
    } else {
        hist_tile<ROOT, WIDE, G, 0>(a, t, act, sh_hist);
    }
    __syncthreads();
    uint32_t* out = a.d_tile_counts + (size_t)blockIdx.x * NB;
    for (int b = threadIdx.x; b < NB; b += blockDim.x) out[b] = sh_hist[b];
}

// ------------------------------------------------------------------------------------------------
// scan: per digit, exclusive prefix over the tiles of each active node
// ------------------------------------------------------------------------------------------------
__global__ void k_scan_chunk_sums(const __grid_constant__ PassArgs a) {
    const ChunkDesc c = a.d_chunks[blockIdx.x];
    for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) {
        uint32_t s = 0;
        const uint32_t* p = a.d_tile_counts + (size_t)c.tile_begin * a.nbins + b;
        for (uint32_t t = 0; t < c.ntiles; ++t) s += p[(size_t)t * a.nbins];
        a.d_chunk_sums[(size_t)blockIdx.x * a.nbins + b] = s;
    }
}
__global__ void k_scan_nodes(const __grid_constant__ PassArgs a) {
    const ActiveDesc act = a.d_active[blockIdx.x];
    for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) {
        uint64_t run = 0;
        uint32_t* p = a.d_chunk_sums + (size_t)act.chunk_begin * a.nbins + b;
        for (uint32_t c = 0; c < act.nchunks; ++c) {
            uint32_t v = p[(size_t)c * a.nbins];
            p[(size_t)c * a.nbins] = (uint32_t)run;
            run += v;
        }
        a.d_node_bins[(size_t)blockIdx.x * a.nbins + b] = run;
    }
}
__global__ void k_scan_tiles(const __grid_constant__ PassArgs a) {
    const ChunkDesc c = a.d_chunks[blockIdx.x];
    for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) {
        uint32_t run = a.d_chunk_sums[(size_t)blockIdx.x * a.nbins + b];
        uint32_t* p = a.d_tile_counts + (size_t)c.tile_begin * a.nbins + b;
        for (uint32_t t = 0; t < c.ntiles; ++t) {
            uint32_t v = p[(size_t)t * a.nbins];
            p[(size_t)t * a.nbins] = run;
            run += v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// scatter
// ------------------------------------------------------------------------------------------------
// One block (16 warps) per 4096-point tile, warp w owns the contiguous items [256 w, 256 w + 256).
//   stage    the tile's input records and colours are contiguous in HBM: one elected thread issues two TMA bulk copies
//            (cp.async.bulk global -> shared, completion on an mbarrier) that overlap the bucket-table set-up
//   sweep A  descent of every item (2 sub-rounds of 32 items in flight per warp) from the staged record; the kept codes
//            replace the record's codes in shared memory (index and colour pass through); destination bucket staged
//            next to it; per-warp bucket counts with __match_any_sync
//   scan     per-bucket totals, exclusive prefix over the buckets (sorted start inside the tile) and over the warps
//   sweep B1 every warp walks its items in order and writes (item, bucket) at the item's stable sorted position: the tile
//            is sorted by bucket as a permutation in shared memory
//   sweep B2 consecutive threads store consecutive records of a bucket's run to the next pass segment or the leaf arena:
//            a warp's store covers a few contiguous runs instead of up to 32 scattered 16-byte slots (and pages)
// The order inside every bucket is the tile order, i.e. input order (stable).
// 16 warps per tile with two sub-rounds in flight per warp (64 registers, two blocks per SM for narrow records).  Measured
// at N = 1e9: 32 warps with one sub-round in flight (32 registers, 2048 threads per SM) 51.9 ms against 50.4 ms - unlike
// k_hist, the scatter gains nothing from occupancy (its sweeps are bound by shared-memory traffic and barriers).
template <bool WIDE>
struct ScatterCfg {
    static constexpr int threads = 512;
    static constexpr int warps = threads / 32;
    static constexpr int warp_items = kTilePoints / warps;  // 256
    static constexpr int sub_rounds = warp_items / 32;      // 8
    static constexpr int U = 2;                             // sub-rounds in flight per warp
    static_assert(sub_rounds % U == 0, "sub-rounds must be a multiple of the unroll");
};
constexpr size_t kRgbStage = (size_t)kTilePoints * 3 + 32;  // root pass: the tile's rgb bytes, staged with 16-byte loads in the (not yet used) sinfo area
static_assert(kRgbStage <= (size_t)kTilePoints * 4, "the rgb staging area aliases sinfo");

template <bool WIDE>
struct ScatterSmem {
    static constexpr size_t rec_bytes = WIDE ? 32 : 16;
    __host__ __device__ static constexpr size_t bytes(int nb) {
        return (size_t)kTilePoints * rec_bytes + ((size_t)kTilePoints + 4) * 4 + (size_t)kTilePoints * 4 + (size_t)nb * 4 * (2 + ScatterCfg<WIDE>::warps) +
               (size_t)nb * 4 + 16;
    }
};

static_assert(ScatterSmem<false>::bytes(512) <= 232448 && ScatterSmem<true>::bytes(512) <= 232448, "scatter tile exceeds the 227 KB opt-in shared memory of sm_100");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    uint64_t state;
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 %0, [%1], %2;" : "=l"(state) : "r"(smem_u32(bar)), "r"(bytes) : "memory");
    (void)state;
}
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

template <bool WIDE>
__device__ __forceinline__ void smem_load_rec(const unsigned char* srec, uint32_t i, uint64_t c[3], uint32_t& idx) {
    if (WIDE) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(srec) + 2 * (size_t)i;
        const ulonglong2 v0 = p[0], v1 = p[1];
        c[0] = v0.x, c[1] = v0.y, c[2] = v1.x, idx = (uint32_t)v1.y;
    } else {
        const uint4 v = reinterpret_cast<const uint4*>(srec)[i];
        c[0] = v.x, c[1] = v.y, c[2] = v.z, idx = v.w;
    }
}
template <bool WIDE>
__device__ __forceinline__ void smem_store_rec(unsigned char* srec, uint32_t i, const uint64_t c[3], uint32_t idx) {
    if (WIDE) {
        ulonglong2* p = reinterpret_cast<ulonglong2*>(srec) + 2 * (size_t)i;
        p[0] = make_ulonglong2(c[0], c[1]);
        p[1] = make_ulonglong2(c[2], (unsigned long long)idx);
    } else {
        reinterpret_cast<uint4*>(srec)[i] = make_uint4((uint32_t)c[0], (uint32_t)c[1], (uint32_t)c[2], idx);
    }
}

template <bool ROOT, bool WIDE, int G, int FAST, typename CodeT>
__device__ __forceinline__ unsigned scatter_sweep_a(const PassArgs& a, const TileDesc& t, const ActiveDesc& act, int warp, int lane, unsigned char* srec,
                                                    uint32_t* sinfo, const uint32_t* lutm, uint32_t* cnt) {
    constexpr int nb = 1 << (3 * G);
    constexpr int kWarpItems = ScatterCfg<WIDE>::warp_items, kSubRounds = ScatterCfg<WIDE>::sub_rounds, kScatterU = ScatterCfg<WIDE>::U;
    unsigned bad = 0;
    for (int s0 = 0; s0 < kSubRounds; s0 += kScatterU) {
        double q[kScatterU][3], m[kScatterU][3];
        uint32_t idx[kScatterU];
        unsigned bin[kScatterU];
        CodeT cj[kScatterU][G][3];
#pragma unroll
        for (int u = 0; u < kScatterU; ++u) {
            const uint32_t i = warp * kWarpItems + (s0 + u) * 32 + lane;
            const bool valid = i < t.count;
            if (ROOT) {
                const uint32_t ic = min(i, t.count - 1);
                load_position_t<true, WIDE, ENC_F64>(a, t, act, ic, q[u], idx[u]);
                if (FAST == 2) bad |= input_bad(q[u][0]) | input_bad(q[u][1]) | input_bad(q[u][2]);
            } else {
                uint64_t c[3] = {0, 0, 0};  // lanes past the end of the tile descend from the cube's min corner (harmless)
                idx[u] = 0;
                if (valid) smem_load_rec<WIDE>(srec, i, c, idx[u]);
                PCV_ENC_SWITCH(a.lv.enc[a.level], _Pragma("unroll") for (int k = 0; k < 3; ++k) q[u][k] = decode_axis<ENC>(c[k], act.m[k], act.e);)
            }
            m[u][0] = act.m[0];
            m[u][1] = act.m[1];
            m[u][2] = act.m[2];
            bin[u] = 0;
        }
        double e = act.e;
#pragma unroll
        for (int j = 1; j <= G; ++j) {
            const double eh = a.lv.edge[a.level + j], ry = a.lv.ry[a.level + j];
            if (j < G) {
                PCV_ENC_SWITCH(a.lv.enc[a.level + j], _Pragma("unroll") for (int u = 0; u < kScatterU; ++u) {
                    bin[u] = (bin[u] << 3) | level_step<ENC, FAST, true>(q[u], m[u], e, eh, ry, cj[u][j - 1], bad);
                })
            } else {  // last level: the decoded position is not needed any more
                PCV_ENC_SWITCH(a.lv.enc[a.level + j], _Pragma("unroll") for (int u = 0; u < kScatterU; ++u) {
                    bin[u] = (bin[u] << 3) | level_step<ENC, FAST, false>(q[u], m[u], e, eh, ry, cj[u][j - 1], bad);
                })
            }
            e = eh;
        }
#pragma unroll
        for (int u = 0; u < kScatterU; ++u) {
            const uint32_t i = warp * kWarpItems + (s0 + u) * 32 + lane;
            const uint32_t lm = lutm[bin[u]];  // local bucket | keep << 16
            const uint32_t lbv = i < t.count ? (lm & 0xFFFFu) : 0xFFFFu;
            const int keep = (lm >> 16) & 0xFF;
            uint64_t c[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                CodeT v = cj[u][0][k];
                if (G >= 2 && keep == 2) v = cj[u][G >= 2 ? 1 : 0][k];
                if (G >= 3 && keep == 3) v = cj[u][G >= 3 ? 2 : 0][k];
                c[k] = (uint64_t)v;
            }
            if (lbv != 0xFFFFu) smem_store_rec<WIDE>(srec, i, c, idx[u]);
            // group the sub-round's lanes by bucket once; sweep B reuses rank / leader / group size
            const unsigned mask = __match_any_sync(0xffffffffu, lbv);
            const uint32_t leader = (uint32_t)__ffs(mask) - 1u, gsize = (uint32_t)__popc(mask), rank = (uint32_t)__popc(mask & ((1u << lane) - 1u));
            sinfo[i] = lbv | (rank << 16) | (leader << 21) | ((gsize - 1u) << 26);
            if (lbv != 0xFFFFu && lane == (int)leader) atomicAdd(&cnt[warp * nb + lbv], gsize);
        }
    }
    return bad;
}

template <bool ROOT, bool WIDE, int G>
__global__ void __launch_bounds__(ScatterCfg<WIDE>::threads, WIDE ? 1 : 2) k_scatter(const __grid_constant__ PassArgs a) {
    typedef typename std::conditional<WIDE, uint64_t, uint32_t>::type CodeT;
    constexpr int nb = 1 << (3 * G);
    constexpr int kScatterThreads = ScatterCfg<WIDE>::threads, kScatterWarps = ScatterCfg<WIDE>::warps, kWarpItems = ScatterCfg<WIDE>::warp_items,
                  kSubRounds = ScatterCfg<WIDE>::sub_rounds;
    constexpr size_t recsz = ScatterSmem<WIDE>::rec_bytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* srec = smem_raw;                                                      // [tile] records (AoS, as in HBM)
    uint32_t* scol_base = reinterpret_cast<uint32_t*>(srec + (size_t)kTilePoints * recsz);  // [tile + 4] colours
    uint2* bdst = reinterpret_cast<uint2*>(scol_base + kTilePoints + 4);                 // [nb] {first slot of this tile, sorted start | leaf << 31}
    uint32_t* cnt = reinterpret_cast<uint32_t*>(bdst + nb);                              // [warps][nb]
    uint32_t* sinfo = cnt + kScatterWarps * nb;                                          // [tile] bucket | rank | leader | group size | leaf
    uint32_t* lutm = sinfo + kTilePoints;                                                // [nb] digit -> bucket | keep << 16 | leaf << 24
    uint8_t* srgb = reinterpret_cast<uint8_t*>(sinfo);                                   // root pass only: rgb bytes of the tile, consumed before sweep A writes sinfo
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + ((ScatterSmem<WIDE>::bytes(nb) - 8) & ~(size_t)7));

    const TileDesc t = tile_of(a, blockIdx.x);
    const ActiveDesc act = a.d_active[t.active];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // (0) TMA: bulk-copy the tile's records and colours into shared memory while the tables below are set up
    uint32_t* scol = scol_base;
    if (!ROOT) {
        const uint32_t coff = (uint32_t)(t.start & 3);  // the colour source must be 16-byte aligned: start 0..3 entries early
        scol = scol_base + coff;
        if (tid == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t rec_bytes = t.count * (uint32_t)recsz;
            const uint32_t col_bytes = ((coff + t.count) * 4u + 15u) & ~15u;
            mbar_expect_tx(mbar, rec_bytes + col_bytes);
            tma_bulk_load(srec, reinterpret_cast<const unsigned char*>(a.rec_in) + t.start * recsz, rec_bytes, mbar);
            tma_bulk_load(scol_base, a.col_in + (t.start - coff), col_bytes, mbar);
        }
    }

    // (1) exclusive prefix of this node's earlier tiles, per digit -> inclusive scan over digits
    const uint32_t* pfx = a.d_tile_counts + (size_t)blockIdx.x * nb;
    for (int b = tid; b < nb; b += kScatterThreads) {
        cnt[b] = pfx[b];
        lutm[b] = 0xFFFFu;  // digits without points keep an invalid bucket
    }
    __syncthreads();
    if (warp == 0) {
        const int per = (nb + 31) / 32;
        const int b0 = lane * per, b1 = min(nb, b0 + per);
        uint32_t s = 0;
        for (int b = b0; b < b1; ++b) s += cnt[b];
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        uint32_t run = incl - s;
        for (int b = b0; b < b1; ++b) {
            run += cnt[b];
            cnt[b] = run;  // inclusive over digits
        }
    }
    __syncthreads();
    // (2) per local bucket: destination of this tile's first record
    for (int lb = tid; lb < nb; lb += kScatterThreads) {
        const BucketDesc bd = a.d_buckets[(size_t)t.active * nb + lb];
        uint32_t v = 0;
        if (bd.b1 != 0) {
            const uint32_t hi = cnt[bd.b1 - 1], lo = bd.b0 ? cnt[bd.b0 - 1] : 0u;
            v = (uint32_t)bd.dest + (hi - lo);
            // every digit of the bucket's range learns (bucket, keep, leaf) in one table entry
            for (uint32_t d = bd.b0; d < bd.b1; ++d) lutm[d] = (uint32_t)lb | ((uint32_t)bd.keep << 16);
        }
        bdst[lb] = make_uint2(v, bd.b1 != 0 && bd.kind ? 0x80000000u : 0u);
    }
    __syncthreads();
    for (int i = tid; i < kScatterWarps * nb; i += kScatterThreads) cnt[i] = 0;
    if (ROOT) {
        // stage the tile's rgb bytes with aligned 16-byte loads (3-byte-strided per-thread loads are LSU-hostile)
        const uint8_t* g0 = a.pts.rgb + 3 * t.start;
        const uint32_t nbytes = 3 * t.count;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(g0) & 15);  // srgb[mis + k] = g0[k]
        const uint8_t* ga = g0 - mis;
        const uint32_t nvec = (mis + nbytes + 15) / 16;
        for (uint32_t v = tid; v + 1 < nvec; v += kScatterThreads) reinterpret_cast<uint4*>(srgb)[v] = __ldcg(reinterpret_cast<const uint4*>(ga) + v);
        if (tid < 16) {  // the last vector byte-wise: never read past the array's last byte
            const uint32_t k = (nvec - 1) * 16 + tid;
            if (k >= mis && k < mis + nbytes) srgb[k] = __ldg(ga + k);
        }
        __syncthreads();
        // colours of 4 consecutive points are 12 staged bytes: 4 aligned words, funnel-shifted by the misalignment, give
        // the 4 packed colours with one 16-byte store (instead of 3 byte loads and a store per point inside sweep A)
        const uint32_t* w = reinterpret_cast<const uint32_t*>(srgb);
        for (uint32_t k = tid; 4 * k < t.count; k += kScatterThreads) {
            const uint32_t byte0 = mis + 12 * k, wi = byte0 >> 2, sh = (byte0 & 3) * 8;
            const uint32_t w0 = w[wi], w1 = w[wi + 1], w2 = w[wi + 2], w3 = w[wi + 3];
            const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh), a2 = __funnelshift_r(w2, w3, sh);
            uint4 c;  // a0 = r0 g0 b0 r1 | a1 = g1 b1 r2 g2 | a2 = b2 r3 g3 b3 (little endian)
            c.x = a0 & 0xFFFFFFu;
            c.y = (a0 >> 24) | ((a1 & 0xFFFFu) << 8);
            c.z = (a1 >> 16) | ((a2 & 0xFFu) << 16);
            c.w = a2 >> 8;
            reinterpret_cast<uint4*>(scol)[k] = c;
        }
    } else {
        mbar_wait(mbar, 0);  // the staged records / colours have landed
    }
    __syncthreads();

    // (3) sweep A (speculatively through the reciprocal division; redone with the IEEE operator if any numerator of
    // the block was outside the proven range - the codes in shared memory are only replaced by valid lanes, so the
    // second run must start from the original records: reload them)
    if (a.lv.fast) {
        const unsigned bad = a.lv.fast == 2 ? scatter_sweep_a<ROOT, WIDE, G, 2, CodeT>(a, t, act, warp, lane, srec, sinfo, lutm, cnt)
                                            : scatter_sweep_a<ROOT, WIDE, G, 1, CodeT>(a, t, act, warp, lane, srec, sinfo, lutm, cnt);
        if (__syncthreads_or((int)bad)) {
            for (int i = tid; i < kScatterWarps * nb; i += kScatterThreads) cnt[i] = 0;
            if (!ROOT) {  // restore the input records (sweep A overwrote their codes)
                for (uint32_t i = tid; i < t.count; i += kScatterThreads) {
                    uint64_t c[3];
                    uint32_t idx;
                    load_rec<WIDE>(a.rec_in, t.start + i, c, idx);
                    smem_store_rec<WIDE>(srec, i, c, idx);
                }
            }
            __syncthreads();
            scatter_sweep_a<ROOT, WIDE, G, 0, CodeT>(a, t, act, warp, lane, srec, sinfo, lutm, cnt);
            __syncthreads();
        }
    } else {
        scatter_sweep_a<ROOT, WIDE, G, 0, CodeT>(a, t, act, warp, lane, srec, sinfo, lutm, cnt);
        __syncthreads();
    }
    // (4) sort the tile by bucket inside shared memory (as a permutation) so that the global stores are coalesced runs:
    //     per-bucket totals -> exclusive scan over the buckets (sorted start of every bucket) -> per-warp offsets
    for (int lb = tid; lb < nb; lb += kScatterThreads) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < kScatterWarps; ++w) tot += cnt[w * nb + lb];
        lutm[lb] = tot;  // the digit table is dead after sweep A
    }
    __syncthreads();
    if (warp == 0) {
        const int per = (nb + 31) / 32;
        const int b0 = lane * per, b1 = min(nb, b0 + per);
        uint32_t sum = 0;
        for (int b = b0; b < b1; ++b) sum += lutm[b];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        uint32_t run = incl - sum;
        for (int b = b0; b < b1; ++b) {
            const uint32_t c = lutm[b];
            lutm[b] = run;  // exclusive: sorted position of the bucket's first record
            run += c;
        }
    }
    __syncthreads();
    for (int lb = tid; lb < nb; lb += kScatterThreads) {
        uint32_t run = lutm[lb];
        bdst[lb].y |= run;
#pragma unroll
        for (int w = 0; w < kScatterWarps; ++w) {
            const uint32_t c = cnt[w * nb + lb];
            cnt[w * nb + lb] = run;
            run += c;
        }
    }
    // every thread takes the sweep-A notes of its items into registers: the permutation is built in the same array
    uint32_t info[kSubRounds];
#pragma unroll
    for (int s = 0; s < kSubRounds; ++s) info[s] = sinfo[warp * kWarpItems + s * 32 + lane];
    __syncthreads();
    // (5) sweep B1: stable sorted position of every item (warp by warp, sub-round by sub-round, rank inside the group)
    uint32_t* perm = sinfo;  // [tile] item index | bucket << 12, in sorted order
#pragma unroll
    for (int s = 0; s < kSubRounds; ++s) {
        const uint32_t i = warp * kWarpItems + s * 32 + lane;
        const uint32_t lbv = info[s] & 0xFFFFu, rank = (info[s] >> 16) & 31u, gsize = ((info[s] >> 26) & 31u) + 1u;
        const int leader = (int)((info[s] >> 21) & 31u);
        uint32_t old = 0;
        if (lane == leader && lbv != 0xFFFFu) {
            old = cnt[warp * nb + lbv];
            cnt[warp * nb + lbv] = old + gsize;
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        if (lbv != 0xFFFFu) perm[old + rank] = i | (lbv << 12);
        __syncwarp();
    }
    __syncthreads();
    // (6) sweep B2: consecutive threads store consecutive records of a bucket's run (next pass segment or leaf arena)
    for (uint32_t p = tid; p < t.count; p += kScatterThreads) {
        const uint32_t e = perm[p], i = e & (kTilePoints - 1), lb = e >> 12;
        const uint2 bd = bdst[lb];
        const bool leaf = (bd.y >> 31) != 0;
        const uint32_t dst = bd.x + (p - (bd.y & 0x7FFFFFFFu));
        uint64_t c64[3];
        uint32_t idx;
        smem_load_rec<WIDE>(srec, i, c64, idx);
        store_rec<WIDE>(leaf ? a.arena : a.rec_next, dst, c64, idx);
        (leaf ? a.col_arena : a.col_next)[dst] = scol[i];
    }
}

// ------------------------------------------------------------------------------------------------
// place
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_code(uint8_t* p, uint64_t c, int enc) {
    if (enc == ENC_U8)
        *p = (uint8_t)c;
    else if (enc == ENC_U16)
        *reinterpret_cast<uint16_t*>(p) = (uint16_t)c;
    else if (enc == ENC_F32)
        *reinterpret_cast<uint32_t*>(p) = (uint32_t)c;
    else
        *reinterpret_cast<uint64_t*>(p) = c;
}

__device__ __forceinline__ void place_store(const PlaceArgs& a, const DNode& nd, uint64_t slot, const uint64_t c[3], uint32_t col, uint32_t idx) {
    const uint64_t dp = nd.out_point_off + slot;
    const int bpc = enc_bytes(nd.enc);
    uint8_t* px = a.out_xyz + nd.out_xyz_off + slot * 3 * (uint64_t)bpc;
    store_code(px, c[0], nd.enc);
    store_code(px + bpc, c[1], nd.enc);
    store_code(px + 2 * bpc, c[2], nd.enc);
    uint8_t* rd = a.out_rgb + 3ull * dp;
    rd[0] = (uint8_t)col;
    rd[1] = (uint8_t)(col >> 8);
    rd[2] = (uint8_t)(col >> 16);
    a.out_src[dp] = idx;
    if (a.out_intensity) a.out_intensity[dp] = __ldg(a.pts.intensity + idx);
}

// The 7 of 8 points of a non-root node that stay: one same-cube rewrite (child_writer, generation.rs:234-238) with the
// node's encoding hoisted out of the loop; dense lane mapping (stayer s <-> rank j = 8*(s/7) + s%7 + 1).
template <bool WIDE, int ENC, int FAST>
__device__ __forceinline__ unsigned place_stayers(const PlaceArgs& a, const LeafTile& lt, const DNode& nd) {
    unsigned bad = 0;
    const uint32_t nst = lt.count - (lt.count + 7) / 8;  // tile starts at a multiple of 8
    for (uint32_t s = threadIdx.x; s < nst; s += blockDim.x) {
        const uint32_t i = 8 * (s / 7) + (s % 7) + 1;
        uint64_t c[3];
        uint32_t idx;
        load_rec<WIDE>(a.arena, lt.arena_start + i, c, idx);
        const uint32_t col = __ldg(a.col_arena + lt.arena_start + i);
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = encode_axis<ENC, FAST>(decode_axis<ENC>(c[k], nd.m[k], nd.e), nd.m[k], nd.e, nd.ry, bad);
        if (!FAST || !bad) {
            const uint64_t j = lt.j0 + i;
            place_store(a, nd, j - (j >> 3) - 1, c, col, idx);
        }
    }
    return bad;
}

// Points that move (every 8th by current rank, generation.rs:224-238): walk up re-encoding through every cube.
template <bool WIDE>
__device__ __forceinline__ void place_movers(const PlaceArgs& a, const LeafTile& lt, const DNode& leaf, bool all_points) {
    const uint32_t step = all_points ? 1u : 8u;
    for (uint32_t i = threadIdx.x * step; i < lt.count; i += blockDim.x * step) {
        uint64_t c[3];
        uint32_t idx;
        load_rec<WIDE>(a.arena, lt.arena_start + i, c, idx);
        const uint32_t col = __ldg(a.col_arena + lt.arena_start + i);
        uint64_t j = lt.j0 + i;
        DNode nd = leaf;
        while (nd.parent >= 0 && (j & 7) == 0) {
            const DNode P = a.d_nodes[nd.parent];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (a.fast) {
                    const double q = decode1_fast(c[k], nd.m[k], nd.e, nd.enc);
                    c[k] = encode1_fast(q, P.m[k], P.e, P.ry, P.enc);
                } else {
                    const double q = decode1(c[k], nd.m[k], nd.e, nd.enc);
                    c[k] = encode1(q, P.m[k], P.e, P.enc);
                }
            }
            j = nd.off_in_parent + (j >> 3);
            nd = P;
        }
        uint64_t slot = j;
        if (nd.parent >= 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (a.fast) {
                    const double q = decode1_fast(c[k], nd.m[k], nd.e, nd.enc);
                    c[k] = encode1_fast(q, nd.m[k], nd.e, nd.ry, nd.enc);
                } else {
                    const double q = decode1(c[k], nd.m[k], nd.e, nd.enc);
                    c[k] = encode1(q, nd.m[k], nd.e, nd.enc);
                }
            }
            slot = j - (j >> 3) - 1;
        }
        place_store(a, nd, slot, c, col, idx);
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(256) k_place(const __grid_constant__ PlaceArgs a) {
    const LeafTile lt = leaf_tile_of(a, blockIdx.x);
    const DNode leaf = a.d_nodes[lt.node];
    if (threadIdx.x == 32 && a.prefetch_tiles) {  // leaf tiles are consecutive in the arena
        const uint64_t first = lt.arena_start + (uint64_t)a.prefetch_tiles * kPlaceTile;
        if (first < a.npoints) {
            const uint64_t cnt = min((uint64_t)kPlaceTile, a.npoints - first), rb = WIDE ? 32 : 16;
            l2_prefetch(reinterpret_cast<const unsigned char*>(a.arena) + first * rb, cnt * rb);
            l2_prefetch(a.col_arena + first, cnt * 4);
        }
    }
    // A tile whose start rank is not a multiple of 8 (top assembly: a collector's points start at arbitrary ranks), or
    // whose node ends the walk (root / collector), goes through the generic per-point path.
    const bool generic = leaf.parent < 0 || (lt.j0 & 7) != 0;
    if (generic) {
        place_movers<WIDE>(a, lt, leaf, true);
        return;
    }
    if (a.fast) {
        unsigned bad = 0;
        if (a.fast == 2) {
            PCV_ENC_SWITCH(leaf.enc, bad = place_stayers<WIDE, ENC, 2>(a, lt, leaf);)
        } else {
            PCV_ENC_SWITCH(leaf.enc, bad = place_stayers<WIDE, ENC, 1>(a, lt, leaf);)
        }
        if (__syncthreads_or((int)bad)) {  // rare: redo the tile's stayers with the IEEE operator (idempotent stores)
            PCV_ENC_SWITCH(leaf.enc, place_stayers<WIDE, ENC, 0>(a, lt, leaf);)
        }
    } else {
        PCV_ENC_SWITCH(leaf.enc, place_stayers<WIDE, ENC, 0>(a, lt, leaf);)
    }
    place_movers<WIDE>(a, lt, leaf, false);
}

// ------------------------------------------------------------------------------------------------
// CUDA backend
// ------------------------------------------------------------------------------------------------
struct CudaBackend : Backend {
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // optional per-kernel timing (CUDA events on the launching stream around every launch)
    enum { K_BBOX = 0, K_HIST, K_SCAN, K_SCATTER, K_PLACE, K_PLY, K_COUNT };
    struct KStat {
        uint64_t launches = 0, bytes = 0;
        double ms = 0;
    };
    bool profile = false;
    KStat kstat[K_COUNT];
    struct Pending {
        int k;
        cudaEvent_t e0, e1;
        uint64_t bytes;
    };
    std::vector<Pending> pending;
    void prof_begin(int k, uint64_t bytes) {
        if (!profile) return;
        Pending p{k, nullptr, nullptr, bytes};
        cudaEventCreate(&p.e0);
        cudaEventCreate(&p.e1);
        cudaEventRecord(p.e0, stream);
        pending.push_back(p);
    }
    void prof_end() {
        if (!profile) return;
        cudaEventRecord(pending.back().e1, stream);
    }
    void prof_collect() {  // call after a stream synchronize
        for (auto& p : pending) {
            float ms = 0;
            cudaEventElapsedTime(&ms, p.e0, p.e1);
            kstat[p.k].launches++;
            kstat[p.k].ms += ms;
            kstat[p.k].bytes += p.bytes;
            cudaEventDestroy(p.e0);
            cudaEventDestroy(p.e1);
        }
        pending.clear();
    }
    void prof_reset() {
        for (auto& k : kstat) k = KStat();
    }

    explicit CudaBackend(cudaStream_t s) : stream(s) {
        for (auto& e : ev) PCV_CUDA_CHECK(cudaEventCreate(&e));
        allow_smem_all();
    }
    ~CudaBackend() override {
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (pin) cudaFreeHost(pin);
        if (pin_back) cudaFreeHost(pin_back);
    }
    void* dmalloc(size_t bytes) override {
        void* p = nullptr;
        PCV_CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 16, stream));
        return p;
    }
    void dfree(void* p) override {
        if (p) cudaFreeAsync(p, stream);
    }
    // Host -> device staging through a pinned ring so that descriptor uploads are truly asynchronous (a pageable
    // source would make cudaMemcpyAsync wait for all prior work of the stream).  The ring is recycled after a
    // stream synchronize (d2h below, once per pass) or when it wraps.
    uint8_t* pin = nullptr;
    size_t pin_cap = 0, pin_off = 0;
    void h2d(void* d, const void* h, size_t bytes) override {
        if (bytes == 0) return;
        const size_t need = (bytes + 255) & ~(size_t)255;
        if (need > pin_cap / 2) {  // too big for the ring: grow it (rare) after draining the stream
            PCV_CUDA_CHECK(cudaStreamSynchronize(stream));
            if (pin) cudaFreeHost(pin);
            pin_cap = std::max<size_t>(need * 4, (size_t)64 << 20);
            PCV_CUDA_CHECK(cudaMallocHost(&pin, pin_cap));
            pin_off = 0;
        }
        if (pin_off + need > pin_cap) {
            PCV_CUDA_CHECK(cudaStreamSynchronize(stream));
            pin_off = 0;
        }
        std::memcpy(pin + pin_off, h, bytes);
        PCV_CUDA_CHECK(cudaMemcpyAsync(d, pin + pin_off, bytes, cudaMemcpyHostToDevice, stream));
        pin_off += need;
    }
    uint8_t* pin_back = nullptr;
    size_t pin_back_cap = 0;
    void d2h(void* h, const void* d, size_t bytes) override {
        if (bytes > pin_back_cap) {
            if (pin_back) cudaFreeHost(pin_back);
            pin_back_cap = std::max<size_t>(bytes * 2, (size_t)8 << 20);
            PCV_CUDA_CHECK(cudaMallocHost(&pin_back, pin_back_cap));
        }
        PCV_CUDA_CHECK(cudaMemcpyAsync(pin_back, d, bytes, cudaMemcpyDeviceToHost, stream));
        PCV_CUDA_CHECK(cudaStreamSynchronize(stream));
        std::memcpy(h, pin_back, bytes);
        pin_off = 0;  // everything staged before this point has been consumed
    }
    void mark(int what) override { cudaEventRecord(ev[what], stream); }

    template <bool ROOT, bool WIDE, int G>
    static void allow_smem() {
        cudaFuncSetAttribute(k_scatter<ROOT, WIDE, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ScatterSmem<WIDE>::bytes(1 << (3 * G)));
    }
    static void allow_smem_all() {
        allow_smem<false, false, 1>(), allow_smem<false, false, 2>(), allow_smem<false, false, 3>();
        allow_smem<true, false, 1>(), allow_smem<true, false, 2>(), allow_smem<true, false, 3>();
        allow_smem<false, true, 1>(), allow_smem<false, true, 2>(), allow_smem<false, true, 3>();
        allow_smem<true, true, 1>(), allow_smem<true, true, 2>(), allow_smem<true, true, 3>();
    }

    template <bool ROOT, bool WIDE>
    void launch_hist(const PassArgs& a, size_t sm) {
        if (a.G == 1)
            k_hist<ROOT, WIDE, 1><<<a.ntiles, 256, sm, stream>>>(a);
        else if (a.G == 2)
            k_hist<ROOT, WIDE, 2><<<a.ntiles, 256, sm, stream>>>(a);
        else
            k_hist<ROOT, WIDE, 3><<<a.ntiles, 256, sm, stream>>>(a);
    }
    // prefetch distance = blocks resident at once (SMs x blocks per SM); 0 disables (PCV_NO_PREFETCH=1 for experiments)
    int sms = 0;
    bool no_prefetch = false;
    uint32_t resident(int blocks_per_sm) {
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            no_prefetch = std::getenv("PCV_NO_PREFETCH") != nullptr;
        }
        return no_prefetch ? 0u : (uint32_t)(sms * blocks_per_sm);
    }
    void hist(const PassArgs& a) override {
        const size_t sm = (size_t)a.nbins * 4;
        const uint64_t rec = a.wide ? sizeof(RecW) : sizeof(RecN);
        prof_begin(K_HIST, a.npoints * (a.root ? 24 : rec));
        if (a.root) {
            if (a.wide)
                launch_hist<true, true>(a, sm);
            else
                launch_hist<true, false>(a, sm);
        } else {
            if (a.wide)
                launch_hist<false, true>(a, sm);
            else
                launch_hist<false, false>(a, sm);
        }
        prof_end();
        ++launches;
        PCV_CUDA_CHECK(cudaGetLastError());
    }
    void scan(const PassArgs& a) override {
        const int th = a.nbins < 32 ? 32 : (a.nbins > 512 ? 512 : a.nbins);
        prof_begin(K_SCAN, (uint64_t)a.ntiles * a.nbins * 12);
        k_scan_chunk_sums<<<a.nchunks, th, 0, stream>>>(a);
        k_scan_nodes<<<a.nactive, th, 0, stream>>>(a);
        k_scan_tiles<<<a.nchunks, th, 0, stream>>>(a);
        prof_end();
        launches += 3;
        PCV_CUDA_CHECK(cudaGetLastError());
    }
    template <bool ROOT, bool WIDE>
    void launch_scatter(const PassArgs& a, size_t sm) {
        if (a.G == 1)
            k_scatter<ROOT, WIDE, 1><<<a.ntiles, ScatterCfg<WIDE>::threads, sm, stream>>>(a);
        else if (a.G == 2)
            k_scatter<ROOT, WIDE, 2><<<a.ntiles, ScatterCfg<WIDE>::threads, sm, stream>>>(a);
        else
            k_scatter<ROOT, WIDE, 3><<<a.ntiles, ScatterCfg<WIDE>::threads, sm, stream>>>(a);
    }
    void scatter(const PassArgs& a) override {
        const size_t sm = a.wide ? ScatterSmem<true>::bytes(a.nbins) : ScatterSmem<false>::bytes(a.nbins);
        const uint64_t rec = a.wide ? sizeof(RecW) : sizeof(RecN);
        prof_begin(K_SCATTER, a.npoints * ((a.root ? 27 : rec + 4) + rec + 4));
        if (a.root) {
            if (a.wide)
                launch_scatter<true, true>(a, sm);
            else
                launch_scatter<true, false>(a, sm);
        } else {
            if (a.wide)
                launch_scatter<false, true>(a, sm);
            else
                launch_scatter<false, false>(a, sm);
        }
        prof_end();
        ++launches;
        PCV_CUDA_CHECK(cudaGetLastError());
    }
    void place(const PlaceArgs& a_in) override {
        if (a_in.ntiles == 0) return;
        PlaceArgs a = a_in;
        a.prefetch_tiles = resident(4);
        prof_begin(K_PLACE, a.npoints * ((a.wide ? sizeof(RecW) : sizeof(RecN)) + 3 + 3 + 4 + (a.out_intensity ? 8 : 0)) + a.xyz_bytes);
        if (a.wide)
            k_place<true><<<a.ntiles, 256, 0, stream>>>(a);
        else
            k_place<false><<<a.ntiles, 256, 0, stream>>>(a);
        prof_end();
        ++launches;
        PCV_CUDA_CHECK(cudaGetLastError());
    }

    void bbox(const PointsView& p, double mn[3], double mx[3]) {
        if (p.n == 0) {  // generation.rs:269: unwrap_or_else(Aabb::zero)
            for (int a = 0; a < 3; ++a) mn[a] = mx[a] = 0.0;
            return;
        }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int blocks = (int)std::min<uint64_t>((uint64_t)sms * 8, (p.n + 255) / 256);
        double* d = (double*)dmalloc((size_t)blocks * 6 * 8);
        prof_begin(K_BBOX, p.n * 24);
        k_bbox<<<blocks, 256, 0, stream>>>(p, d);
        prof_end();
        ++launches;
        PCV_CUDA_CHECK(cudaGetLastError());
        std::vector<double> h((size_t)blocks * 6);
        d2h(h.data(), d, h.size() * 8);
        dfree(d);
        for (int a = 0; a < 3; ++a) {
            mn[a] = h[a];
            mx[a] = h[3 + a];
        }
        for (int b = 1; b < blocks; ++b)
            for (int a = 0; a < 3; ++a) {
                mn[a] = std::fmin(mn[a], h[(size_t)b * 6 + a]);
                mx[a] = std::fmax(mx[a], h[(size_t)b * 6 + 3 + a]);
            }
    }
};

}  // namespace pcv
