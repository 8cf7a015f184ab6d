// pcv_synth.h — counter-based synthetic point generators for benchmarks and parity tests, bit-identical on host and device.
// Input-data definition only (no part of the octree algorithm): shared by the CUDA library (device generator for bench.py),
// by the oracle's C API (host generator for the CPU reference arm, so that arm never loads the CUDA library) and by tests.
// Only integer hashing, exact int->f64 conversion and single IEEE mul/add steps (no libm, no FMA contraction).
//
//   PCV_SYNTH_SLAB_ECEF       shape of point_cloud_test's SyntheticData (synthetic_data.rs:22-78,
//                             point_cloud_test/src/lib.rs:42-53): 200 x 200 x 20 m slab, uniform, in a
//                             local ENU-like frame placed at ECEF magnitude; RGB = 24-bit point index.
//                             (The reference draws from rand 0.7 ChaCha; that stream cannot be
//                             regenerated here, so the distribution is restated with splitmix64.)
//   PCV_SYNTH_GAUSS_CLUSTERS  BASELINE.json config 2: 4096 Gaussian-like clusters (Irwin-Hall(12)
//                             offsets, sigma in [0.5, 8) m) in a 1024 m cube, resolution 1024/2^20, plus 8
//                             blocks of 150 000 identical points (global indices 2^20 ..) that exercise the
//                             "too small to be split" branch (generation.rs:137-147).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PCV_SYNTH_HD __host__ __device__ __forceinline__
#else
#define PCV_SYNTH_HD inline
#endif

namespace pcv {

PCV_SYNTH_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct SynthFrame {  // fixed rigid transform local -> "ECEF"
    double r[9];
    double t[3];
};
PCV_SYNTH_HD SynthFrame slab_frame() {
    // rotation = Rz(0.7) * Ry(-0.9) rounded to f64 (orthonormal to 1e-16); translation ~ 6.37e6 m from origin
    SynthFrame f;
    f.r[0] = 0.47543352776997644;
    f.r[1] = -0.644217687237691;
    f.r[2] = -0.5991214669182833;
    f.r[3] = 0.4004521361232219;
    f.r[4] = 0.7648421872844885;
    f.r[5] = -0.5046330500712651;
    f.r[6] = 0.7833269096274834;
    f.r[7] = 0.0;
    f.r[8] = 0.6216099682706644;
    f.t[0] = 4157222.543;
    f.t[1] = 664789.307;
    f.t[2] = 4774952.099;
    return f;
}
PCV_SYNTH_HD void frame_apply(const SynthFrame& f, const double l[3], double out[3]) {
    for (int i = 0; i < 3; ++i) {
        double a = f.r[3 * i] * l[0];
        a = a + f.r[3 * i + 1] * l[1];
        a = a + f.r[3 * i + 2] * l[2];
        out[i] = a + f.t[i];
    }
}

constexpr double kGaussEdge = 1024.0;
constexpr uint64_t kDegenerateFirst = 1ull << 20;
constexpr uint64_t kDegenerateBlock = 150000;

PCV_SYNTH_HD void synth_point(int kind, uint64_t seed, uint64_t idx, double p[3], uint8_t rgb[3]) {
    const uint64_t h0 = splitmix64(idx ^ seed);
    if (kind == 1) {  // slab
        const uint64_t h1 = splitmix64(h0), h2 = splitmix64(h1);
        double l[3];
        l[0] = (double)(uint32_t)(h0 >> 32) * (200.0 / 4294967296.0) - 100.0;
        l[1] = (double)(uint32_t)(h1 >> 32) * (200.0 / 4294967296.0) - 100.0;
        l[2] = (double)(uint32_t)(h2 >> 32) * (20.0 / 4294967296.0) - 10.0;
        const SynthFrame f = slab_frame();
        frame_apply(f, l, p);
        rgb[0] = (uint8_t)(idx >> 16);  // synthetic_data.rs:67-73
        rgb[1] = (uint8_t)(idx >> 8);
        rgb[2] = (uint8_t)idx;
        return;
    }
    const double gmin[3] = {300000.125, -200000.5, 1000.25};
    rgb[0] = (uint8_t)(h0 >> 16);
    rgb[1] = (uint8_t)(h0 >> 8);
    rgb[2] = (uint8_t)h0;
    if (idx >= kDegenerateFirst && idx < kDegenerateFirst + 8 * kDegenerateBlock) {
        const uint64_t d = (idx - kDegenerateFirst) / kDegenerateBlock;
        const uint64_t hd = splitmix64(0xD1CEull + d + seed);
        for (int a = 0; a < 3; ++a) p[a] = gmin[a] + (double)((hd >> (20 * a)) & 0xFFFFF) * (1024.0 / 1048576.0);
        return;
    }
    const uint64_t c = h0 >> 52;  // 4096 clusters
    const uint64_t hc = splitmix64(0xC1A57E5ull + c + (seed << 1));
    const double sigma = 0.5 + (double)(uint32_t)(splitmix64(hc) >> 48) * (7.5 / 65536.0);
    uint64_t h = h0;
    for (int a = 0; a < 3; ++a) {
        const double centre = gmin[a] + (double)((hc >> (20 * a)) & 0xFFFFF) * (1024.0 / 1048576.0);
        uint32_t sum = 0;  // Irwin-Hall(12) of u16
        for (int k = 0; k < 3; ++k) {
            h = splitmix64(h);
            sum += (uint32_t)(h & 0xFFFF) + (uint32_t)((h >> 16) & 0xFFFF) + (uint32_t)((h >> 32) & 0xFFFF) + (uint32_t)(h >> 48);
        }
        const double g = ((double)sum - 393210.0) * (1.0 / 65536.0);
        double v = centre + g * sigma;
        const double hi = gmin[a] + 1024.0;
        if (v < gmin[a]) v = gmin[a];
        if (v > hi) v = hi;
        p[a] = v;
    }
}

// SyntheticData::bbox (synthetic_data.rs:46-50) / the cluster cube, and the resolution each workload is quoted with.
inline void synth_bbox(int kind, double bbox_min[3], double bbox_max[3], double* resolution) {
    if (kind == 1) {
        const SynthFrame f = slab_frame();
        for (int i = 0; i < 8; ++i) {
            double l[3] = {(i & 1) ? 100.0 : -100.0, (i & 2) ? 100.0 : -100.0, (i & 4) ? 10.0 : -10.0}, p[3];
            frame_apply(f, l, p);
            for (int a = 0; a < 3; ++a) {
                bbox_min[a] = (i == 0 || p[a] < bbox_min[a]) ? p[a] : bbox_min[a];
                bbox_max[a] = (i == 0 || p[a] > bbox_max[a]) ? p[a] : bbox_max[a];
            }
        }
        if (resolution) *resolution = 0.001;  // point_cloud_test/src/lib.rs:45
        return;
    }
    const double gmin[3] = {300000.125, -200000.5, 1000.25};
    for (int a = 0; a < 3; ++a) {
        bbox_min[a] = gmin[a];
        bbox_max[a] = gmin[a] + 1024.0;
    }
    if (resolution) *resolution = 1024.0 / 1048576.0;
}

}  // namespace pcv
