// lod_order.h — the per-node draw order of the viewers (sdl_viewer/src/node_drawer.rs:34-43,185-205: "We draw the points in
// random order. This allows us to only draw the first N if we want to draw less"; octree/mod.rs:286-287 asks for that order to be
// applied when the node is written).  The reference shuffles with thread_rng at load time, so the order itself is not a parity
// target - only that it is a permutation applied identically to positions and colours.  Here it is a keyed bijection on [0, n)
// that needs no table and no sort: a few invertible mixing rounds on ceil(log2 n) bits, cycle-walking back into range.
#pragma once
#include <stdint.h>

#include "chain.h"

namespace pcv {

PCV_HD uint64_t lod_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
PCV_HD uint64_t lod_node_key(uint64_t seed, uint64_t id_high, uint64_t id_low) { return lod_mix64(lod_mix64(seed ^ id_high) ^ id_low); }

// new_order[i] for a node of n points: position i of the shuffled node holds the old point lod_order(key, n, i).
PCV_HD uint32_t lod_order(uint64_t key, uint32_t n, uint32_t i) {
    if (n <= 1) return 0;
    int bits = 1;
    while (bits < 32 && (1ull << bits) < n) ++bits;
    const uint32_t mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    const uint64_t k1 = lod_mix64(key), k2 = lod_mix64(k1);
    const uint32_t m1 = (uint32_t)k1 | 1u, m2 = (uint32_t)(k1 >> 32) | 1u, a1 = (uint32_t)k2, a2 = (uint32_t)(k2 >> 32);
    const int sh = bits > 1 ? bits / 2 : 1;
    uint32_t x = i;
    do {  // every step is a bijection on `bits`-bit integers; values that fall outside [0, n) walk on (n > 2^(bits-1): < 2 steps on average)
        x = (x * m1 + a1) & mask;
        x ^= x >> sh;
        x = (x * m2 + a2) & mask;
        x ^= x >> sh;
    } while (x >= n);
    return x;
}

}  // namespace pcv
