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

#ifndef KS_STREAM_STORES
#define KS_STREAM_STORES 1  // region stores bypass L2 retention (read once by the next level): -30 us on the C2 rank kernel
#endif

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
// Returns the number of keys of the tile.
// `dst`: DS_TILE keys of shared memory for the bin-ordered copy of the tile (may be the buffer the keys were read from).
__device__ __forceinline__ uint32_t scatter_keys(const uint64_t (&key)[DS_ITEMS], uint32_t valid, const DenseScatter& sc,
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
#if KS_STREAM_STORES
        if (g < sc.cap) __stcs(reinterpret_cast<unsigned long long*>(sc.out) + (uint64_t)(bucket_base + b) * sc.cap + g, (unsigned long long)k);
#else
        if (g < sc.cap) sc.out[(uint64_t)(bucket_base + b) * sc.cap + g] = k;
#endif
    }
    return total;
}

// The same for (hash, loc) pairs of the general path (hashes that rarely repeat: pairs are distinct, their final order
// is numeric, and the bucket sort breaks ties between equal hashes by loc, so no stability is needed here either).
// bin = the `bits` bits of the normalised hash (key << lz) that start `shift` bits from the bottom.
struct PairScatter {
    uint64_t *out_key, *out_val;  // regions of `cap` entries each, region index = bucket_base + bin
    uint32_t* cursor;
    uint32_t cap;
    int shift, bits, lz;
    uint32_t* overflow;
};

// fetch_val(i) returns the value that goes with key[i]; it is called after the keys have been moved, so the values
// may live in a buffer that `dst_val` aliases (the sketch kernel's staging), while `dst_key` may alias the keys' source.
template <class FetchVal>
__device__ __forceinline__ uint32_t scatter_pairs(const uint64_t (&key)[DS_ITEMS], uint32_t valid, FetchVal fetch_val,
                                              const PairScatter& sc, uint32_t bucket_base, DenseScatterSmem& sm,
                                              uint64_t* dst_key, uint64_t* dst_val) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nbins = 1u << sc.bits, mask = nbins - 1u;
    auto bin_of = [&](uint64_t k) -> uint32_t { return sc.bits ? (uint32_t)((k << sc.lz) >> sc.shift) & mask : 0u; };
    if (tid < nbins) sm.hist[tid] = 0;
    __syncthreads();
    uint32_t packed[DS_ITEMS];  // bin << 12 | slot in the bin (a tile holds 2048 pairs)
#pragma unroll
    for (int i = 0; i < DS_ITEMS; i++)
        if ((valid >> i) & 1u) {
            const uint32_t b = bin_of(key[i]);
            packed[i] = (b << 12) | atomicAdd(&sm.hist[b], 1u);
        }
    __syncthreads();
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
        if (tid == DS_THREADS - 1) sm.start[nbins] = off + incl;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < DS_ITEMS; i++)
        if ((valid >> i) & 1u) dst_key[sm.start[packed[i] >> 12] + (packed[i] & 0xfffu)] = key[i];
    uint64_t val[DS_ITEMS];
#pragma unroll
    for (int i = 0; i < DS_ITEMS; i++)
        if ((valid >> i) & 1u) val[i] = fetch_val(i);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < DS_ITEMS; i++)
        if ((valid >> i) & 1u) dst_val[sm.start[packed[i] >> 12] + (packed[i] & 0xfffu)] = val[i];
    __syncthreads();
    const uint32_t total = sm.start[nbins];
    for (uint32_t p = tid; p < total; p += DS_THREADS) {
        const uint64_t k = dst_key[p];
        const uint32_t b = bin_of(k);
        const uint32_t g = sm.gbase[b] + (p - sm.start[b]);
        if (g < sc.cap) {
            const uint64_t at = (uint64_t)(bucket_base + b) * sc.cap + g;
            sc.out_key[at] = k;
            sc.out_val[at] = dst_val[p];
        }
    }
    return total;
}

// The second partition level works on DS_TILE-sized chunks of the first-level regions.  chunk_map[c] says what chunk c is:
// x = region | keys in the chunk << 16 (0 keys: the grid is an upper bound, nothing to do), y = offset of the chunk's first
// key inside its region.  One CTA: scan of the regions' chunk counts, then thread b writes the entries of region b.  The
// partition kernels read ONE broadcast word pair at their head instead of searching a prefix table (eight dependent global
// loads first, then a shared-memory copy + search by every thread: 21 % of dense_partition_kernel's instructions).
static __global__ void dense_chunks_kernel(const uint32_t* __restrict__ cursor1, uint32_t nb1, uint32_t cap1,
                                           uint2* __restrict__ chunk_map, uint32_t max_chunks) {
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cnt = tid < nb1 ? min(cursor1[tid], cap1) : 0u;  // nb1 <= 256
    const uint32_t c = (cnt + DS_TILE - 1) / DS_TILE;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
    for (uint32_t w = 0; w < warp; w++) off += s_w[w];
    const uint32_t first_chunk = off + incl - c;
    if (tid == 255) s_total = off + incl;
    for (uint32_t i = 0; i < c; i++) {
        const uint32_t at = first_chunk + i, first = i * DS_TILE;
        if (at < max_chunks) chunk_map[at] = make_uint2(tid | (min(cnt - first, (uint32_t)DS_TILE) << 16), first);
    }
    __syncthreads();
    for (uint32_t at = s_total + tid; at < max_chunks; at += 256) chunk_map[at] = make_uint2(0u, 0u);
}

// tuple offset of every final bucket (exclusive scan of the clamped cursors), one CTA: every warp owns a contiguous
// segment of the buckets and walks it 32 at a time with coalesced reads -- once to sum it, once (after the 32 segment
// sums are scanned) to write the offsets.  (Scanning 1024 buckets per iteration behind four barriers took 47 us for 2^16
// buckets; a contiguous chunk per THREAD made every warp load touch 32 sectors and took twice that.)
static __global__ void __launch_bounds__(1024)
dense_bucket_offsets_kernel(const uint32_t* __restrict__ cursor2, uint32_t nb, uint32_t cap, uint32_t* __restrict__ bstart) {
    __shared__ uint32_t s_w[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg = ((nb + 31) / 32 + 31) & ~31u;  // a multiple of 32: segments start on a 128-byte line
    const uint32_t b0 = min(warp * seg, nb), b1 = min(b0 + seg, nb);
    uint32_t sum = 0;
#pragma unroll 4
    for (uint32_t i = b0 + lane; i < b1; i += 32) sum += min(cursor2[i], cap);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_w[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_w[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if ((int)lane >= o) wi += t;
        }
        s_w[lane] = wi - w;
        if (lane == 31) bstart[nb] = wi;
    }
    __syncthreads();
    uint32_t run = s_w[warp];
    for (uint32_t base = b0; base < b1; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = i < b1 ? min(cursor2[i], cap) : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (i < b1) bstart[i] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

}  // namespace ks
