// s2.cuh — kernels of the S2-cell point cloud: per-point cell keys with the S2Splitter's validity rule, the
// gather into cell-contiguous arrays, the CellUnion point test.  The arithmetic is csrc/s2.h (shared with the sequential
// test backend); the kernels only distribute points over threads.  The stable grouping by cell between the two is a
// library radix sort (cub::DeviceRadixSort::SortPairs, stable): an HBM-bound pass structure of 8 + 4 bytes per point and
// digit, not worth a second hand-written partition next to the octree's.
#pragma once
#include <cuda_runtime.h>

#include "build_host.hpp"
#include "s2.h"

namespace pcv {

// keys[i] = CellID::from_point(p_i).parent(level), idx[i] = i; *first_bad = the smallest index of a point that "is not a
// valid ECEF point" (read_write/s2.rs:64-71), or ~0.
__global__ void __launch_bounds__(256) k_s2_keys(const PointsView p, int level, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                                 unsigned long long* __restrict__ first_bad) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double x = p.x[i * p.stride], y = p.y[i * p.stride], z = p.z[i * p.stride];
        if (first_bad && !s2_valid_ecef(x, y, z)) atomicMin(first_bad, (unsigned long long)i);
        keys[i] = s2_parent(s2_cell_id_from_point(x, y, z), level);
        if (idx) idx[i] = (uint32_t)i;
    }
}

// number of positions whose key differs from its predecessor (= number of cells in a sorted key array)
__global__ void __launch_bounds__(256) k_s2_count_runs(const uint64_t* __restrict__ keys, uint64_t n, unsigned long long* __restrict__ runs) {
    unsigned int local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        local += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(runs, (unsigned long long)local);
}

// slot s of the cloud holds input point order[s]: positions as f64 x, y, z triples (Encoding::Plain), colours, intensity.
struct S2GatherArgs {
    PointsView p;
    const uint32_t* order;
    double* xyz;       // n * 3
    uint8_t* rgb;      // n * 3 or null
    float* intensity;  // n or null
};
__global__ void __launch_bounds__(256) k_s2_gather(const S2GatherArgs a) {
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < a.p.n; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = a.order[s];
        a.xyz[3 * s + 0] = a.p.x[i * a.p.stride];
        a.xyz[3 * s + 1] = a.p.y[i * a.p.stride];
        a.xyz[3 * s + 2] = a.p.z[i * a.p.stride];
        if (a.rgb) {
            a.rgb[3 * s + 0] = a.p.rgb[3 * i + 0];
            a.rgb[3 * s + 1] = a.p.rgb[3 * i + 1];
            a.rgb[3 * s + 2] = a.p.rgb[3 * i + 2];
        }
        if (a.intensity) a.intensity[s] = a.p.intensity[i];
    }
}

// CellUnion as a PointCulling (geometry/s2_cell_union.rs:27-31): flag[k] = union.contains_cellid(CellID::from_point(p_k)).
// `cells` is the normalised union (sorted); points are AoS triples (xyz) or a PointsView.
__global__ void __launch_bounds__(256) k_s2_union_mask(const PointsView p, uint64_t first, uint64_t count, const uint64_t* __restrict__ cells, uint32_t ncells,
                                                       uint8_t* __restrict__ flag) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + k;
        const uint64_t id = s2_cell_id_from_point(p.x[i * p.stride], p.y[i * p.stride], p.z[i * p.stride]);
        flag[k] = s2_union_contains(cells, ncells, id) ? 1 : 0;
    }
}

// survivors of a query: slot list -> output arrays
struct S2EmitArgs {
    const uint64_t* slots;  // selected slots (ascending)
    uint64_t n;
    const double* xyz;
    const uint8_t* rgb;
    const float* intensity;
    const uint32_t* src;
    double* xyz_out;
    uint8_t* rgb_out;
    float* intensity_out;
    uint64_t* src_out;
};
__global__ void __launch_bounds__(256) k_s2_emit(const S2EmitArgs a) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.n; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = a.slots[k];
        a.xyz_out[3 * k + 0] = a.xyz[3 * s + 0];
        a.xyz_out[3 * k + 1] = a.xyz[3 * s + 1];
        a.xyz_out[3 * k + 2] = a.xyz[3 * s + 2];
        if (a.rgb_out) {
            a.rgb_out[3 * k + 0] = a.rgb[3 * s + 0];
            a.rgb_out[3 * k + 1] = a.rgb[3 * s + 1];
            a.rgb_out[3 * k + 2] = a.rgb[3 * s + 2];
        }
        if (a.intensity_out) a.intensity_out[k] = a.intensity[s];
        if (a.src_out) a.src_out[k] = a.src[s];
    }
}

}  // namespace pcv
