// kernels_shard.cuh — multi-GPU sharding helpers: level-k prefix cell of every point (the first k steps
// of the re-quantising descent on the raw positions, i.e. exactly the cell the single-GPU build would route the
// point to), its 8^k histogram, and a stable pack of the points by destination rank into contiguous send buffers.
#pragma once
#include <cuda_runtime.h>

#include "kernels_build.cuh"

namespace pcv {

struct PrefixArgs {
    PointsView pts;
    LevelTable lv;
    double root_min[3];
    int k;
    int nbins;  // 8^k
};

__device__ __forceinline__ unsigned prefix_cell(const PrefixArgs& a, uint64_t i) {
    double q[3] = {__ldg(a.pts.x + i * a.pts.stride), __ldg(a.pts.y + i * a.pts.stride), __ldg(a.pts.z + i * a.pts.stride)};
    double m[3] = {a.root_min[0], a.root_min[1], a.root_min[2]};
    double e = a.lv.edge[0];
    unsigned cell = 0;
    for (int j = 1; j <= a.k; ++j) {
        Step s = descend_any(a.lv, j, q, m, e);
        cell = (cell << 3) | s.digit;
        e = a.lv.edge[j];
    }
    return cell;
}

__global__ void __launch_bounds__(256) k_prefix_hist(const __grid_constant__ PrefixArgs a, unsigned long long* __restrict__ counts) {
    extern __shared__ uint32_t sh_cnt[];
    for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) sh_cnt[b] = 0;
    __syncthreads();
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.pts.n; i += step) atomicAdd(&sh_cnt[prefix_cell(a, i)], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < a.nbins; b += blockDim.x)
        if (sh_cnt[b]) atomicAdd(&counts[b], (unsigned long long)sh_cnt[b]);
}

// ---- pack -------------------------------------------------------------------------------------------------------
constexpr uint32_t kPackTile = 4096;
constexpr int kMaxRanks = 64;

struct PackArgs {
    PrefixArgs p;
    const int32_t* cell_to_rank;  // device, nbins
    uint32_t nranks, ntiles;
    uint8_t* dest;        // per point destination rank (written by the count pass)
    uint32_t* counts;     // [nranks][ntiles]; after the scan: first output slot of (rank, tile)
    const uint64_t* gidx_in;  // optional global indices of the local points
    uint64_t gidx_base;       // used when gidx_in == nullptr: global index = base + i
    double* out_xyz;          // n * 3 (AoS)
    uint8_t* out_rgb;
    float* out_intensity;
    uint64_t* out_idx;
};

__global__ void __launch_bounds__(256) k_pack_count(const __grid_constant__ PackArgs a) {
    __shared__ uint32_t cnt[kMaxRanks];
    if (threadIdx.x < kMaxRanks) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t t0 = (uint64_t)blockIdx.x * kPackTile;
    const uint32_t n = (uint32_t)min((uint64_t)kPackTile, a.p.pts.n - t0);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = a.cell_to_rank[prefix_cell(a.p, t0 + i)];
        a.dest[t0 + i] = (uint8_t)r;
        atomicAdd(&cnt[r], 1u);
    }
    __syncthreads();
    if (threadIdx.x < a.nranks) a.counts[(size_t)threadIdx.x * a.ntiles + blockIdx.x] = cnt[threadIdx.x];
}

__global__ void __launch_bounds__(256) k_pack_scatter(const __grid_constant__ PackArgs a) {
    __shared__ uint32_t base[kMaxRanks];
    __shared__ uint32_t wc[8][kMaxRanks];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < a.nranks) base[threadIdx.x] = a.counts[(size_t)threadIdx.x * a.ntiles + blockIdx.x];
    const uint64_t t0 = (uint64_t)blockIdx.x * kPackTile;
    const uint32_t n = (uint32_t)min((uint64_t)kPackTile, a.p.pts.n - t0);
    for (uint32_t r0 = 0; r0 < n; r0 += 256) {
        for (int i = threadIdx.x; i < 8 * kMaxRanks; i += 256) (&wc[0][0])[i] = 0;
        __syncthreads();
        const uint32_t i = r0 + threadIdx.x;
        const uint32_t d = i < n ? a.dest[t0 + i] : 0xFFu;
        const unsigned mask = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        if (d != 0xFFu && lane == __ffs(mask) - 1) wc[warp][d] = __popc(mask);
        __syncthreads();
        uint32_t before = 0;
        if (d != 0xFFu) {
            for (int w = 0; w < warp; ++w) before += wc[w][d];
            const uint64_t dst = (uint64_t)base[d] + before + rank;
            const uint64_t g = t0 + i;
            a.out_xyz[3 * dst] = __ldg(a.p.pts.x + g * a.p.pts.stride);
            a.out_xyz[3 * dst + 1] = __ldg(a.p.pts.y + g * a.p.pts.stride);
            a.out_xyz[3 * dst + 2] = __ldg(a.p.pts.z + g * a.p.pts.stride);
            a.out_rgb[3 * dst] = __ldg(a.p.pts.rgb + 3 * g);
            a.out_rgb[3 * dst + 1] = __ldg(a.p.pts.rgb + 3 * g + 1);
            a.out_rgb[3 * dst + 2] = __ldg(a.p.pts.rgb + 3 * g + 2);
            if (a.out_intensity) a.out_intensity[dst] = __ldg(a.p.pts.intensity + g);
            a.out_idx[dst] = a.gidx_in ? __ldg(a.gidx_in + g) : a.gidx_base + g;
        }
        __syncthreads();
        if (threadIdx.x < a.nranks) {
            uint32_t s = 0;
            for (int w = 0; w < 8; ++w) s += wc[w][threadIdx.x];
            base[threadIdx.x] += s;
        }
        __syncthreads();
    }
}

}  // namespace pcv
