// Shared device helpers for the kmerseek B200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

constexpr uint64_t C1 = 0x87c37b91114253d5ULL;
constexpr uint64_t C2 = 0x4cf5ad432745937fULL;
constexpr uint64_t SEED = 42;

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
__host__ __device__ __forceinline__ uint64_t mix_k1(uint64_t k1) { return rotl64(k1 * C1, 31) * C2; }
__host__ __device__ __forceinline__ uint64_t mix_k2(uint64_t k2) { return rotl64(k2 * C2, 33) * C1; }

// MurmurHash3_x64_128 (seed 42, low word) of k bytes, one byte at a time: the generic-k sketch kernel (k > 32) and the
// query kernel (queries are small) use it; the templated limb version of sketch.cu is the fast path of the build.
__device__ __forceinline__ uint64_t murmur_bytes(const uint8_t* s, uint32_t k) {
    uint64_t h1 = SEED, h2 = SEED;
    uint32_t nb = k / 16;
    for (uint32_t b = 0; b < nb; b++) {
        uint64_t k1 = 0, k2 = 0;
        for (int i = 0; i < 8; i++) {
            k1 |= (uint64_t)s[16 * b + i] << (8 * i);
            k2 |= (uint64_t)s[16 * b + 8 + i] << (8 * i);
        }
        h1 ^= mix_k1(k1);
        h1 = rotl64(h1, 27) + h2;
        h1 = h1 * 5 + 0x52dce729;
        h2 ^= mix_k2(k2);
        h2 = rotl64(h2, 31) + h1;
        h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t* t = s + 16 * nb;
    uint32_t rem = k & 15;
    uint64_t k1 = 0, k2 = 0;
    for (uint32_t i = 8; i < rem; i++) k2 |= (uint64_t)t[i] << (8 * (i - 8));
    for (uint32_t i = 0; i < (rem < 8 ? rem : 8); i++) k1 |= (uint64_t)t[i] << (8 * i);
    if (rem > 8) h2 ^= mix_k2(k2);
    if (rem > 0) h1 ^= mix_k1(k1);
    h1 ^= k;
    h2 ^= k;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    return h1 + h2;
}

// ---------------------------------------------------------------------------------------------
// Single-pass chained scan state ("decoupled look-back").  One 64-bit word per tile carries the
// flag in the top two bits and the value in the low 62, so a reader never sees a flag without its
// value and no fence is needed between them.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t SCAN_AGG = 1ULL << 62;  // tile aggregate available
constexpr uint64_t SCAN_PFX = 2ULL << 62;  // inclusive prefix available
constexpr uint64_t SCAN_VAL = (1ULL << 62) - 1;

__device__ __forceinline__ uint64_t ld_relaxed(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ONE full warp of the tile.  Publishes this tile's aggregate, walks back over the
// predecessors' words 64 at a time (2 per lane, independent loads), returns the exclusive prefix (same
// value in every lane) and publishes the inclusive prefix.  `value` must be < 2^62.
// The window is wide on purpose: with hundreds of tiles in flight none of a tile's near predecessors has
// a prefix yet, and every 32-wide round would cost a full L2 round trip.
// First half of scan_lookback: make this tile's aggregate visible (tile 0: its inclusive prefix).  A tile that
// has other work left can publish early and collect its prefix later with scan_collect.
__device__ __forceinline__ void scan_publish(uint64_t* status, uint32_t tile, uint64_t value) {
    if ((threadIdx.x & 31) == 0) st_relaxed(status + tile, (tile == 0 ? SCAN_PFX : SCAN_AGG) | value);
}

__device__ __forceinline__ uint64_t scan_collect(uint64_t* status, uint32_t tile, uint64_t value);

__device__ __forceinline__ uint64_t scan_lookback(uint64_t* status, uint32_t tile, uint64_t value) {
    scan_publish(status, tile, value);
    return scan_collect(status, tile, value);
}

// Second half: walk back over the predecessors, return the exclusive prefix, publish the inclusive one.
__device__ __forceinline__ uint64_t scan_collect(uint64_t* status, uint32_t tile, uint64_t value) {
    constexpr int W = 2;
    const uint32_t lane = threadIdx.x & 31;
    if (tile == 0) return 0;
    uint64_t excl = 0;
    int64_t base = (int64_t)tile - 1;
    while (true) {
        // lane l looks at predecessors base - (l*W + j), j = 0..W-1: distance grows with (lane, j)
        uint64_t w[W];
        bool ready;
        bool first_poll = true;
        do {
            if (!first_poll) __nanosleep(100);  // leave the issue slots to the warps that still hash
            first_poll = false;
            ready = true;
#pragma unroll
            for (int j = 0; j < W; j++) {
                int64_t idx = base - (int64_t)(lane * W + j);
                w[j] = idx >= 0 ? ld_relaxed(status + idx) : SCAN_PFX;  // tiles before 0: prefix 0
                ready &= (w[j] >> 62) != 0;
            }
            // only the words nearer than the nearest prefix matter; older tiles publish earlier, so waiting
            // for the whole window costs nothing extra and keeps the loop simple
        } while (!__all_sync(0xffffffffu, ready));
        // nearest prefix: smallest distance d = lane*W + j with flag == 2
        int my_first = W;  // j of this lane's nearest prefix word
#pragma unroll
        for (int j = W - 1; j >= 0; j--) if ((w[j] >> 62) == 2) my_first = j;
        const uint32_t pfx_mask = __ballot_sync(0xffffffffu, my_first < W);
        uint64_t v = 0;
        if (pfx_mask) {
            const int first_lane = __ffs(pfx_mask) - 1;
            if ((int)lane < first_lane) {
#pragma unroll
                for (int j = 0; j < W; j++) v += w[j] & SCAN_VAL;
            } else if ((int)lane == first_lane) {
#pragma unroll
                for (int j = 0; j < W; j++) if (j <= my_first) v += w[j] & SCAN_VAL;
            }
        } else {
#pragma unroll
            for (int j = 0; j < W; j++) v += w[j] & SCAN_VAL;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (pfx_mask) break;
        base -= 32 * W;
    }
    if (lane == 0) st_relaxed(status + tile, SCAN_PFX | ((excl + value) & SCAN_VAL));
    return excl;
}

}  // namespace ks
