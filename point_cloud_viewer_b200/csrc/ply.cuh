// ply.cuh — raw PLY vertex records (AoS, arbitrary property mix, byte-aligned) -> the SoA arrays of the octree build, on
// the GPU.  Restates what PlyIterator::next + batch_from_readers do per point (src/read_write/ply.rs:453-556):
//   position = (x as f64, y as f64, z as f64) + header offset;  colour = (r, g, b) uchar;  intensity = f32
// and folds find_bounding_box (src/octree/generation.rs:256-270) into the same pass.
//
// HBM-bound byte work.  One block per tile of records: the tile is one contiguous byte range, so a single TMA bulk copy
// (cp.async.bulk + mbarrier) stages it in shared memory; threads then pick their record's fields with byte-granular
// shared-memory loads (records are not aligned to anything), write x / y / z / intensity coalesced, and the 3-byte
// colours go through a shared staging area so that they leave as 16-byte vectors.
// Algorithmic bytes per point: record_bytes read + 24 (+3 colour, +4 intensity) written.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pcv.h"
#include "kernels_build.cuh"  // mbarrier / bulk-copy helpers, warp_min / warp_max

namespace pcv {

struct PlyUnpackArgs {
    const uint8_t* raw;  // records of this launch, 16-byte aligned
    uint64_t n;          // records in this launch
    uint64_t out_first;  // index of record 0 in the output arrays (multiple of tile_points)
    uint32_t record_bytes, tile_points;
    int32_t type[3];
    uint32_t off[3], off_rgb[3], off_intensity;
    int32_t has_color, has_intensity;
    double offset[3];
    double *x, *y, *z;
    uint8_t* rgb;
    float* intensity;
    double* partial;  // [gridDim.x][6] min xyz, max xyz of the block's points
};

constexpr int kPlyThreads = 256;

__host__ __device__ inline uint32_t ply_tile_points(uint32_t record_bytes) {
    uint32_t t = (65536u / record_bytes) & ~255u;  // <= 64 KB of records per tile, whole multiples of the block size
    return t > 4096u ? 4096u : (t < 256u ? 256u : t);
}
__host__ __device__ inline size_t ply_smem_bytes(uint32_t record_bytes, uint32_t tile_points) {
    return (((size_t)tile_points * record_bytes + 15) & ~(size_t)15) + 16 + (size_t)tile_points * 3 + 32 + 16;
}

__device__ __forceinline__ uint32_t ld_le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t ld_le32(const uint8_t* p, bool aligned) {
    if (aligned) return *reinterpret_cast<const uint32_t*>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
// `$reading_fn(buf) as f64` (ply.rs:248-284).  int8 is read through `buf[0]`, i.e. as an unsigned byte (ply.rs:254).
__device__ __forceinline__ double ply_as_f64(const uint8_t* p, int type, bool aligned) {
    switch (type) {
        case PCV_PLY_U8:
        case PCV_PLY_I8: return (double)p[0];
        case PCV_PLY_U16: return (double)ld_le16(p);
        case PCV_PLY_I16: return (double)(int16_t)ld_le16(p);
        case PCV_PLY_U32: return (double)ld_le32(p, aligned);
        case PCV_PLY_I32: return (double)(int32_t)ld_le32(p, aligned);
        case PCV_PLY_F32: return (double)__uint_as_float(ld_le32(p, aligned));
        default: {
            const uint64_t lo = ld_le32(p, aligned), hi = ld_le32(p + 4, aligned);
            return __longlong_as_double((long long)(lo | (hi << 32)));
        }
    }
}

__global__ void __launch_bounds__(kPlyThreads) k_ply_unpack(const __grid_constant__ PlyUnpackArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t T = a.tile_points;
    const size_t rec_area = (((size_t)T * a.record_bytes + 15) & ~(size_t)15) + 16;
    uint8_t* srec = smem_raw;
    uint8_t* srgb = smem_raw + rec_area;  // [3 T + 32]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + ((ply_smem_bytes(a.record_bytes, T) - 8) & ~(size_t)7));
    const int tid = threadIdx.x;
    const uint64_t first = (uint64_t)blockIdx.x * T;
    const uint32_t count = (uint32_t)min((uint64_t)T, a.n - first);
    const uint8_t* src = a.raw + first * a.record_bytes;  // 16-byte aligned: T is a multiple of 16
    const uint32_t bytes = count * a.record_bytes, bulk = bytes & ~15u;

    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0 && bulk) {
        mbar_expect_tx(mbar, bulk);
        tma_bulk_load(srec, src, bulk, mbar);
    }
    if (tid < (int)(bytes - bulk)) srec[bulk + tid] = src[bulk + tid];  // the last < 16 bytes of the launch
    if (bulk) mbar_wait(mbar, 0);
    __syncthreads();

    const bool al = ((a.record_bytes | a.off[0] | a.off[1] | a.off[2] | (a.has_intensity ? a.off_intensity : 0u)) & 3u) == 0;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const uint64_t o0 = a.out_first + first;
    for (uint32_t i = tid; i < count; i += kPlyThreads) {
        const uint8_t* p = srec + (size_t)i * a.record_bytes;
        const double x = ply_as_f64(p + a.off[0], a.type[0], al) + a.offset[0];
        const double y = ply_as_f64(p + a.off[1], a.type[1], al) + a.offset[1];
        const double z = ply_as_f64(p + a.off[2], a.type[2], al) + a.offset[2];
        a.x[o0 + i] = x;
        a.y[o0 + i] = y;
        a.z[o0 + i] = z;
        mn[0] = fmin(mn[0], x), mx[0] = fmax(mx[0], x);
        mn[1] = fmin(mn[1], y), mx[1] = fmax(mx[1], y);
        mn[2] = fmin(mn[2], z), mx[2] = fmax(mx[2], z);
        if (a.has_color && a.rgb) {
            srgb[3 * i] = p[a.off_rgb[0]];
            srgb[3 * i + 1] = p[a.off_rgb[1]];
            srgb[3 * i + 2] = p[a.off_rgb[2]];
        }
        if (a.has_intensity && a.intensity) a.intensity[o0 + i] = __uint_as_float(ld_le32(p + a.off_intensity, al));
    }
    __syncthreads();
    if (a.has_color && a.rgb) {  // 3 * o0 is a multiple of 16 (o0 is a multiple of the tile size)
        uint8_t* dst = a.rgb + 3 * o0;
        const uint32_t nb = 3 * count, nvec = nb / 16;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (uint32_t v = tid; v < nvec; v += kPlyThreads) reinterpret_cast<uint4*>(dst)[v] = reinterpret_cast<const uint4*>(srgb)[v];
            for (uint32_t k = nvec * 16 + tid; k < nb; k += kPlyThreads) dst[k] = srgb[k];
        } else {
            for (uint32_t k = tid; k < nb; k += kPlyThreads) dst[k] = srgb[k];
        }
    }
    // block reduction of the bounding box (Aabb::grow = component-wise min / max)
    __shared__ double sh[kPlyThreads / 32][6];
    const int w = tid >> 5, l = tid & 31;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double lo = warp_min(mn[k]), hi = warp_max(mx[k]);
        if (l == 0) {
            sh[w][k] = lo;
            sh[w][3 + k] = hi;
        }
    }
    __syncthreads();
    if (tid < 6) {
        double v = sh[0][tid];
        for (int k = 1; k < kPlyThreads / 32; ++k) v = tid < 3 ? fmin(v, sh[k][tid]) : fmax(v, sh[k][tid]);
        a.partial[(size_t)blockIdx.x * 6 + tid] = v;
    }
}

}  // namespace pcv
