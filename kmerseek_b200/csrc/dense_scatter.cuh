// One level of the dense path's key partition, shared by the rank kernel (sketch.cu: first level, fused into the
// kernel that produces the keys) and dense_partition_kernel (dense.cu: second level).
//
// The keys of the dense path (rank | protein | position) are all distinct and their final order is their numeric
// order, so a partition pass does not have to be stable: a tile's keys are grouped by bin in shared memory, every
// non-empty bin reserves its run in the bin's fixed-capacity output region with ONE global atomic, and the runs go
// out coalesced.  No global histogram, no look-back chain.  A region that overflows (ranks are spread by the hash,
// so only heavy repeats of one k-mer can do that) raises `overflow`; the host then builds the batch on the general path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

constexpr int DS_THREADS = 256;
constexpr int DS_ITEMS = 8;                 // keys per thread
constexpr int DS_TILE = DS_THREADS * DS_ITEMS;
constexpr int DS_MAX_BITS = 8;              // bins per level <= 256

struct DenseScatter {
    uint64_t* out;       // regions of `cap` keys each, region index = bucket_base + bin
    uint32_t* cursor;    // keys already placed in each region
    uint32_t cap;
    int shift, bits;     // bin = (key >> shift) & (2^bits - 1)
    uint32_t* overflow;  // device flag
};

struct DenseScatterSmem {
    uint32_t hist[1 << DS_MAX_BITS];
    uint32_t start[(1 << DS_MAX_BITS) + 1];
    uint32_t gbase[1 << DS_MAX_BITS];
    uint32_t wsum[DS_THREADS / 32];
};

// Called by all DS_THREADS threads.  key[i] is meaningful where bit i of `valid` is set.  Contains block-wide barriers.
// `dst`: DS_TILE keys of shared memory for the bin-ordered copy of the tile (may be the buffer the keys were read from).
__device__ __forceinline__ void scatter_keys(const uint64_t (&key)[DS_ITEMS], uint32_t valid, const DenseScatter& sc,
                                             uint32_t bucket_base, DenseScatterSmem& sm, uint64_t* dst) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nbins = 1u << sc.bits, mask = nbins - 1u;
    if (tid < nbins) sm.hist[tid] = 0;
    __syncthreads();
    uint32_t slot[DS_ITEMS];
#pragma unroll
    for (int i = 0; i < DS_ITEMS; i++)
        if ((valid >> i) & 1u) slot[i] = atomicAdd(&sm.hist[(uint32_t)(key[i] >> sc.shift) & mask], 1u);
    __syncthreads();
    // exclusive scan of the bin counts (thread t owns bin t), one reservation per non-empty bin
    {
        const uint32_t c = tid < nbins ? sm.hist[tid] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        if (lane == 31) sm.wsum[warp] = incl;
        __syncthreads();
        uint32_t off = 0;
#pragma unroll
        for (int w = 0; w < DS_THREADS / 32; w++) off += w < (int)warp ? sm.wsum[w] : 0u;
        if (tid < nbins) {
            sm.start[tid] = off + incl - c;
            uint32_t g = 0;
            if (c) {
                g = atomicAdd(&sc.cursor[bucket_base + tid], c);
                if (g + c > sc.cap) atomicOr(sc.overflow, 1u);
            }
            sm.gbase[tid] = g;
        }
        if (tid == DS_THREADS - 1) sm.start[nbins] = off + incl;  // total (nbins <= DS_THREADS)
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < DS_ITEMS; i++)
        if ((valid >> i) & 1u) dst[sm.start[(uint32_t)(key[i] >> sc.shift) & mask] + slot[i]] = key[i];
    __syncthreads();
    const uint32_t total = sm.start[nbins];
    for (uint32_t p = tid; p < total; p += DS_THREADS) {
        const uint64_t k = dst[p];
        const uint32_t b = (uint32_t)(k >> sc.shift) & mask;
        const uint32_t g = sm.gbase[b] + (p - sm.start[b]);
        if (g < sc.cap) sc.out[(uint64_t)(bucket_base + b) * sc.cap + g] = k;
    }
}

}  // namespace ks
