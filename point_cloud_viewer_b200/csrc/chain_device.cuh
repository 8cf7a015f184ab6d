// chain_device.cuh — the descent step of chain.h specialised for the GPU kernels: the encoding of a level is a
// template parameter (it is uniform for a whole pass level, so the switch is hoisted out of the per-point code) and
// the three axes of a point, and several points per thread, are straight-line code without branches, so their
// independent dependency chains interleave (the kernels are instruction-latency bound, not FP64-throughput bound).
//
// Results are bit-identical to chain.h's descend()/descend_fast() (same operations in the same order); the
// reciprocal-based division is used speculatively and every numerator is range-checked with integer instructions:
// if any check fails the caller repeats the work with the IEEE operator (FAST = false).
#pragma once
#include "chain.h"

namespace pcv {

template <int K>
__device__ __forceinline__ double unit_frac_dev(uint32_t v) {
    return unit_frac<K>(v);
}

// numerator admissible for div_known's fast sequence?  (exponent in [2^-500, 2^500) or exactly +0)
__device__ __forceinline__ unsigned div_bad(double a) {
    const uint32_t hi = (uint32_t)__double2hiint(a), lo = (uint32_t)__double2loint(a);
    const bool inrange = ((hi & 0x7fffffffu) - 0x20b00000u) < 0x3e800000u;
    return (inrange || (hi | lo) == 0u) ? 0u : 1u;
}

// Raw input coordinate admissible for the unchecked fast sequence (FAST == 2)?  |x| < 2^400, which also rejects inf / NaN.
__device__ __forceinline__ unsigned input_bad(double x) { return fabs(x) < 2.582249878086908589655919172003011874329705792829223512830659e120 ? 0u : 1u; }

// FAST: 3 = every edge of the tree is a power of two (LevelTable::fast == 3): a / e is the exact scaling a * 2^-j for EVERY
// input (zero, subnormal, inf, NaN included: both are the correctly rounded value of the same real number), no guard at all;
// 0 = IEEE division; 1 = reciprocal sequence, every numerator range-checked (flag -> the caller redoes with 0);
// 2 = reciprocal sequence without per-numerator checks: the host proved that every cube of the tree keeps all
// coordinates at magnitudes in [2^-300, 2^301) (LevelTable::fast == 2), so q - m is exactly 0 or in [2^-353, 2^500)
// once the raw inputs passed input_bad() in the root pass.
// One axis, one level.  Returns the code; updates q (if DECODE), m, digit bit, bad flag.
template <int ENC, int FAST, bool DECODE>
__device__ __forceinline__ uint64_t axis_step(double& q, double& m, double e_cur, double e_half, double ry, unsigned& bit, unsigned& bad) {
    const double c = (m + (m + e_cur)) * 0.5;  // Cube::center: (min + max) / 2 (halving is exact either way)
    bit = q > c ? 1u : 0u;
    m = bit ? m + e_half : m;
    const double a = q - m;
    double t;
    if (FAST == 3) {
        t = a * ry;
    } else if (FAST) {
        if (FAST == 1) bad |= div_bad(a);
        const double q0 = a * ry;
        const double q1 = fma(fma(-q0, e_half, a), ry, q0);
        t = fma(fma(-q1, e_half, a), ry, q1);
    } else {
        t = a / e_half;
    }
    if (ENC == ENC_U8) {
        uint32_t v = trunc_u32(255.0 * t);  // saturating, NaN -> 0: equals the clamped form (chain.h encode1_fast)
        v = v > 255u ? 255u : v;
        if (DECODE) q = fma(unit_frac<8>(v), e_half, m);
        return v;
    }
    if (ENC == ENC_U16) {
        uint32_t v = trunc_u32(65535.0 * t);
        v = v > 65535u ? 65535u : v;
        if (DECODE) q = fma(unit_frac<16>(v), e_half, m);
        return v;
    }
    const double cl = clamp01(t);
    if (ENC == ENC_F32) {
        const float f = (float)cl;
        if (DECODE) q = fma((double)f, e_half, m);
        return (uint64_t)__float_as_uint(f);
    }
    if (DECODE) q = fma(cl, e_half, m);
    return (uint64_t)__double_as_longlong(cl);
}

// One level for one point: digit = (x > cx) << 2 | (y > cy) << 1 | (z > cz)  (node.rs:34-42)
template <int ENC, int FAST, bool DECODE, typename CodeT>
__device__ __forceinline__ unsigned level_step(double q[3], double m[3], double e_cur, double e_half, double ry, CodeT code[3], unsigned& bad) {
    unsigned bx, by, bz;
    code[0] = (CodeT)axis_step<ENC, FAST, DECODE>(q[0], m[0], e_cur, e_half, ry, bx, bad);
    code[1] = (CodeT)axis_step<ENC, FAST, DECODE>(q[1], m[1], e_cur, e_half, ry, by, bad);
    code[2] = (CodeT)axis_step<ENC, FAST, DECODE>(q[2], m[2], e_cur, e_half, ry, bz, bad);
    return (bx << 2) | (by << 1) | bz;
}

// Only the child digit (the last level of a histogram pass needs no codes).
__device__ __forceinline__ unsigned level_digit(const double q[3], const double m[3], double e_cur) {
    unsigned d = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double c = (m[a] + (m[a] + e_cur)) * 0.5;
        d = (d << 1) | (q[a] > c ? 1u : 0u);
    }
    return d;
}

// decode of a stored code (node file -> position), encoding as template parameter
template <int ENC>
__device__ __forceinline__ double decode_axis(uint64_t bits, double mn, double edge) {
    if (ENC == ENC_U8) return fma(unit_frac<8>((uint32_t)bits), edge, mn);
    if (ENC == ENC_U16) return fma(unit_frac<16>((uint32_t)bits), edge, mn);
    if (ENC == ENC_F32) return fma((double)__uint_as_float((uint32_t)bits), edge, mn);
    return fma(__longlong_as_double((long long)bits), edge, mn);
}

// encode into a known cube (ry = RN(1/edge)), encoding as template parameter; same value as chain.h encode1()
template <int ENC, int FAST>
__device__ __forceinline__ uint64_t encode_axis(double value, double mn, double edge, double ry, unsigned& bad) {
    const double a = value - mn;
    double t;
    if (FAST == 3) {
        t = a * ry;
    } else if (FAST) {
        if (FAST == 1) bad |= div_bad(a);
        const double q0 = a * ry;
        const double q1 = fma(fma(-q0, edge, a), ry, q0);
        t = fma(fma(-q1, edge, a), ry, q1);
    } else {
        t = a / edge;
    }
    if (ENC == ENC_U8) {
        const uint32_t v = trunc_u32(255.0 * t);
        return v > 255u ? 255u : v;
    }
    if (ENC == ENC_U16) {
        const uint32_t v = trunc_u32(65535.0 * t);
        return v > 65535u ? 65535u : v;
    }
    const double cl = clamp01(t);
    if (ENC == ENC_F32) return (uint64_t)__float_as_uint((float)cl);
    return (uint64_t)__double_as_longlong(cl);
}

#define PCV_ENC_SWITCH(enc_value, ...)                  \
    switch (enc_value) {                                \
        case ENC_U8: {                                  \
            constexpr int ENC = ENC_U8;                 \
            __VA_ARGS__                                 \
        } break;                                        \
        case ENC_U16: {                                 \
            constexpr int ENC = ENC_U16;                \
            __VA_ARGS__                                 \
        } break;                                        \
        case ENC_F32: {                                 \
            constexpr int ENC = ENC_F32;                \
            __VA_ARGS__                                 \
        } break;                                        \
        default: {                                      \
            constexpr int ENC = ENC_F64;                \
            __VA_ARGS__                                 \
        } break;                                        \
    }

}  // namespace pcv
