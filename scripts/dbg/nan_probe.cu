// Diagnostic: float -> integer conversions of NaN / inf / huge values at run time (no constant folding) on this GPU.
#include <cstdio>
#include <cmath>
__global__ void k(const double* in, int n, unsigned* u, int* s, unsigned* f) {
    for (int i = 0; i < n; ++i) {
        u[i] = __double2uint_rz(in[i]);
        s[i] = __double2int_rz(in[i]);
        f[i] = __float_as_uint((float)in[i]);
    }
}
int main() {
    double* in; unsigned* u; int* s; unsigned* f;
    const double vals[] = {NAN, -NAN, INFINITY, -INFINITY, 1e300, -1e300, 65535.9, 4294967296.0, -0.5, 2147483648.0, 3e9};
    const int n = sizeof(vals) / sizeof(double);
    cudaMallocManaged(&in, n * 8); cudaMallocManaged(&u, n * 4); cudaMallocManaged(&s, n * 4); cudaMallocManaged(&f, n * 4);
    for (int i = 0; i < n; ++i) in[i] = vals[i];
    k<<<1, 1>>>(in, n, u, s, f);
    cudaDeviceSynchronize();
    for (int i = 0; i < n; ++i) printf("%-12g  u32 0x%08x  s32 0x%08x (%d)  f32bits 0x%08x\n", vals[i], u[i], (unsigned)s[i], s[i], f[i]);
    return 0;
}
