// synth.cuh — device kernel over the shared synthetic point generators (include/pcv_synth.h).
#pragma once
#include "../../include/pcv_synth.h"

namespace pcv {

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) k_synth(int kind, uint64_t seed, uint64_t first, uint64_t n, double* x, double* y, double* z,
                                               uint8_t* rgb) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        double p[3];
        uint8_t c[3];
        synth_point(kind, seed, first + i, p, c);
        x[i] = p[0];
        y[i] = p[1];
        z[i] = p[2];
        rgb[3 * i] = c[0];
        rgb[3 * i + 1] = c[1];
        rgb[3 * i + 2] = c[2];
    }
}
#endif

}  // namespace pcv
