// Dense k-mer space index build for sm_100a: rank tables and CSR construction from sorted rank keys.
//
// For the two-letter hp alphabet a k-mer is a k-bit pattern; with k <= 24 all 2^k patterns can be hashed once per
// handle.  Sorting those hashes gives every pattern its rank, and a (hash, protein, position) tuple collapses into one
// 64-bit key  rank | protein | position.  Keys leave the sketch kernel in (protein, position) order; a stable library
// sort on the rank bits alone puts them in (hash, protein, position) order -- equal ranks are equal hashes, so there is
// no bucket sort -- and two streaming passes turn the sorted keys into the same CSR arrays the general path produces
// (replaces, like index_build.cu, src/rust/index.rs:770-780 and :824-827 of the reference on the hot path).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "dense.cuh"
#include "dense_scatter.cuh"
#include "index_build.cuh"
#include "sketch.cuh"

namespace ks {

namespace {

// MurmurHash3_x64_128 (seed 42, low word) of the k bytes 'h' / 'p' spelled by the low k bits of `pattern` (bit j set =
// byte j is 'h'); k <= 24.  Same arithmetic as the sketch kernels, on whole 64-bit words.
__device__ uint64_t murmur_of_pattern(uint32_t pattern, uint32_t k) {
    uint64_t w[3] = {0, 0, 0};
    for (uint32_t j = 0; j < k; j++) w[j >> 3] |= (uint64_t)(((pattern >> j) & 1u) ? 'h' : 'p') << (8 * (j & 7));
    uint64_t h1 = SEED, h2 = SEED;
    const uint32_t nb = k / 16, rem = k % 16;
    if (nb) {
        h1 ^= mix_k1(w[0]);
        h1 = rotl64(h1, 27) + h2;
        h1 = h1 * 5 + 0x52dce729;
        h2 ^= mix_k2(w[1]);
        h2 = rotl64(h2, 31) + h1;
        h2 = h2 * 5 + 0x38495ab5;
    }
    const uint64_t t1 = w[2 * nb], t2 = nb ? 0 : w[1];  // tail words (k <= 24: a second tail word only without a block)
    if (rem > 8) h2 ^= mix_k2(t2);
    if (rem > 0) h1 ^= mix_k1(t1);
    h1 ^= k;
    h2 ^= k;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    return h1 + h2;
}

// MurmurHash3 of the translated k-mer at residue g of the batch, from the residues themselves (exception keys carry no
// pattern): the byte path of the sketch kernels, one residue at a time -- only exception keys come here.
// hp translation of one residue byte (fill_lut, moltype 2; src/rust/encoding.rs:43-53): hydrophobic AFGILMPVWY -> 'h',
// polar CDEHKNQRST -> 'p', '*' stays, everything else 'X'
__device__ __forceinline__ uint32_t hp_translate(uint32_t c) {
    constexpr uint32_t H = (1u << 0) | (1u << 5) | (1u << 6) | (1u << 8) | (1u << 11) | (1u << 12) | (1u << 15) | (1u << 21) |
                           (1u << 22) | (1u << 24);
    constexpr uint32_t P = (1u << 2) | (1u << 3) | (1u << 4) | (1u << 7) | (1u << 10) | (1u << 13) | (1u << 16) | (1u << 17) |
                           (1u << 18) | (1u << 19);
    if (c == '*') return '*';
    const uint32_t i = c - 'A';
    if (i < 26u) {
        if ((H >> i) & 1u) return 'h';
        if ((P >> i) & 1u) return 'p';
    }
    return 'X';
}

__device__ __noinline__ uint64_t dense_window_hash(const uint8_t* __restrict__ res, int packed, uint64_t g, uint32_t k) {
    uint64_t w[3] = {0, 0, 0};
    for (uint32_t j = 0; j < k; j++) {
        const uint64_t i = g + j;
        uint32_t c;
        if (packed) {
            const uint64_t byte = (i >> 3) * 5;
            uint64_t v = 0;
            for (int t = 0; t < 5; t++) v |= (uint64_t)res[byte + t] << (8 * t);
            const uint32_t code = (uint32_t)(v >> (5 * (i & 7))) & 31u;
            c = code == 0 ? 0u : code <= 26 ? 'A' + code - 1 : code == 27 ? (uint32_t)'*' : 0u;
        } else {
            c = res[i];
        }
        w[j >> 3] |= (uint64_t)hp_translate(c) << (8 * (j & 7));
    }
    uint64_t h1 = SEED, h2 = SEED;
    const uint32_t nb = k / 16, rem = k % 16;
    if (nb) {
        h1 ^= mix_k1(w[0]);
        h1 = rotl64(h1, 27) + h2;
        h1 = h1 * 5 + 0x52dce729;
        h2 ^= mix_k2(w[1]);
        h2 = rotl64(h2, 31) + h1;
        h2 = h2 * 5 + 0x38495ab5;
    }
    const uint64_t t1 = w[2 * nb], t2 = nb ? 0 : w[1];
    if (rem > 8) h2 ^= mix_k2(t2);
    if (rem > 0) h1 ^= mix_k1(t1);
    h1 ^= k;
    h2 ^= k;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    return h1 + h2;
}

__global__ void dense_hash_codes_kernel(uint32_t k, uint64_t* __restrict__ hash, uint32_t* __restrict__ code) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (1u << k)) return;
    hash[c] = murmur_of_pattern(c, k);
    code[c] = c;
}

// group_base[p] = first entry of sorted_hash whose top DENSE_PREFIX_BITS bits are >= p (p = 2^16: n)
__global__ void dense_group_base_kernel(uint32_t n, const uint64_t* __restrict__ sorted_hash, uint32_t* __restrict__ group_base) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > (1u << DENSE_PREFIX_BITS)) return;
    if (p == (1u << DENSE_PREFIX_BITS)) { group_base[p] = n; return; }
    const uint64_t lim = (uint64_t)p << (64 - DENSE_PREFIX_BITS);
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted_hash[mid] < lim) lo = mid + 1; else hi = mid;
    }
    group_base[p] = lo;
}

// flags[0] |= 1 when two patterns share a hash or one hashes to 0; flags[3] = largest rank of a pattern inside its group
__global__ void dense_check_kernel(uint32_t n, const uint64_t* __restrict__ sorted_hash, const uint32_t* __restrict__ group_base,
                                   uint32_t* __restrict__ flags) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t h = sorted_hash[r];
    if (h == 0 || (r && sorted_hash[r - 1] == h)) atomicOr(flags, 1u);
    atomicMax(flags + 3, r - group_base[(uint32_t)(h >> (64 - DENSE_PREFIX_BITS))]);
}

// code_of_pattern[pattern] = prefix << rb | (2 x rank in the prefix group + 1)  (parity layout), or
//                            prefix << rb | rank in the prefix group              (no room for the parity bit):
// see DenseSketchArgs (sketch.cuh)
__global__ void dense_code_kernel(uint32_t n, const uint64_t* __restrict__ sorted_hash, const uint32_t* __restrict__ pattern_of_rank,
                                  const uint32_t* __restrict__ group_base, int rb, int parity, uint32_t* __restrict__ code_of_pattern) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t pfx = (uint32_t)(sorted_hash[r] >> (64 - DENSE_PREFIX_BITS));
    const uint32_t local = r - group_base[pfx];
    code_of_pattern[pattern_of_rank[r]] = (pfx << rb) | (parity ? 2u * local + 1u : local);
}

// every complete window is a tuple (scaled == 1): kept windows per protein; distinct hashes start from the same number
// and lose one for every repeated (hash, protein) pair met while the groups are written
__global__ void dense_windows_kernel(const uint64_t* __restrict__ offsets, uint32_t n_prot, uint32_t k,
                                     uint32_t* __restrict__ t_abund, uint32_t* __restrict__ t_size) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_prot) return;
    const uint64_t len = offsets[p + 1] - offsets[p];
    const uint32_t w = len >= k ? (uint32_t)(len - k + 1) : 0u;
    t_abund[p] = w;
    t_size[p] = w;
}

// The pattern hash behind an odd code: prefix group base + rank in the group.
__device__ __forceinline__ uint64_t hash_of_code(uint32_t code, int rb, int parity, const uint32_t* __restrict__ group_base,
                                                 const uint64_t* __restrict__ sorted_hash) {
    return sorted_hash[group_base[code >> rb] + ((code & ((1u << rb) - 1u)) >> parity)];
}

constexpr int DC_THREADS = 256;
constexpr int DC_ROWS = 8;  // rows of 32 consecutive keys per warp
constexpr int DC_TILE = DC_THREADS * DC_ROWS;

// heads of a tile of sorted keys: a key head where the rank changes, a group head where (rank, protein) changes
__global__ void __launch_bounds__(DC_THREADS)
dense_count_kernel(const uint64_t* __restrict__ keys, uint64_t n, int loc_bits, int pos_bits, uint64_t* __restrict__ tile_counts) {
    __shared__ uint32_t s_k[DC_THREADS / 32], s_g[DC_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * DC_TILE;
    uint32_t ck = 0, cg = 0;
#pragma unroll
    for (int r = 0; r < DC_ROWS; r++) {
        const uint64_t i = base + (uint64_t)(warp * DC_ROWS + r) * 32 + lane;
        bool hk = false, hg = false;
        const uint64_t key = i < n ? keys[i] : 0ull;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (i < n) {
            if (lane == 0) prev = i ? keys[i - 1] : ~key;
            hk = (key >> loc_bits) != (prev >> loc_bits);
            hg = (key >> pos_bits) != (prev >> pos_bits);
        }
        ck += __popc(__ballot_sync(0xffffffffu, hk));
        cg += __popc(__ballot_sync(0xffffffffu, hg));
    }
    if (lane == 0) { s_k[warp] = ck; s_g[warp] = cg; }
    __syncthreads();
    if (tid == 0) {
        uint32_t k = 0, g = 0;
        for (int w = 0; w < DC_THREADS / 32; w++) { k += s_k[w]; g += s_g[w]; }
        tile_counts[blockIdx.x] = (uint64_t)k | ((uint64_t)g << 32);
        if (blockIdx.x == gridDim.x - 1) tile_counts[gridDim.x] = 0;
    }
}

__global__ void __launch_bounds__(DC_THREADS)
dense_write_kernel(const uint64_t* __restrict__ keys_in, uint64_t n, int loc_bits, int pos_bits, const uint64_t* __restrict__ tile_prefix,
                   const uint64_t* __restrict__ sorted_hash, const uint32_t* __restrict__ group_base, int rb, int parity,
                   uint64_t* __restrict__ loc, uint64_t* __restrict__ keys,
                   uint32_t* __restrict__ key_grp, uint32_t* __restrict__ grp_start, uint32_t* __restrict__ t_size,
                   uint64_t* __restrict__ d_counts) {
    __shared__ uint32_t s_k[DC_THREADS / 32], s_g[DC_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t base = (uint64_t)blockIdx.x * DC_TILE;
    const uint64_t pos_mask = (1ull << pos_bits) - 1ull, pid_mask = (1ull << (loc_bits - pos_bits)) - 1ull;
    uint64_t key[DC_ROWS];
    uint32_t bk[DC_ROWS], bg[DC_ROWS];
    uint32_t wk = 0, wg = 0;
#pragma unroll
    for (int r = 0; r < DC_ROWS; r++) {
        const uint64_t i = base + (uint64_t)(warp * DC_ROWS + r) * 32 + lane;
        bool hk = false, hg = false;
        key[r] = i < n ? keys_in[i] : 0ull;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key[r], 1);
        if (i < n) {
            if (lane == 0) prev = i ? keys_in[i - 1] : ~key[r];
            hk = (key[r] >> loc_bits) != (prev >> loc_bits);
            hg = (key[r] >> pos_bits) != (prev >> pos_bits);
            const uint32_t pid = (uint32_t)((key[r] >> pos_bits) & pid_mask);
            loc[i] = ((uint64_t)pid << 32) | (key[r] & pos_mask);
            if (!hg) atomicSub(&t_size[pid], 1u);  // the hash again in the same protein: one distinct hash fewer
        }
        bk[r] = __ballot_sync(0xffffffffu, hk);
        bg[r] = __ballot_sync(0xffffffffu, hg);
        wk += __popc(bk[r]);
        wg += __popc(bg[r]);
    }
    if (lane == 0) { s_k[warp] = wk; s_g[warp] = wg; }
    __syncthreads();
    const uint64_t pfx = tile_prefix[blockIdx.x];
    uint32_t rk = (uint32_t)pfx, rg = (uint32_t)(pfx >> 32);
    uint32_t tk = 0, tg = 0;
#pragma unroll
    for (int w = 0; w < DC_THREADS / 32; w++) {
        if (w < (int)warp) { rk += s_k[w]; rg += s_g[w]; }
        tk += s_k[w];
        tg += s_g[w];
    }
#pragma unroll
    for (int r = 0; r < DC_ROWS; r++) {
        const uint64_t i = base + (uint64_t)(warp * DC_ROWS + r) * 32 + lane;
        const uint32_t g = rg + __popc(bg[r] & lt);
        if ((bg[r] >> lane) & 1u) grp_start[g] = (uint32_t)i;
        if ((bk[r] >> lane) & 1u) {
            const uint32_t u = rk + __popc(bk[r] & lt);
            keys[u] = hash_of_code((uint32_t)(key[r] >> loc_bits), rb, parity, group_base, sorted_hash);  // (no exception keys here)
            key_grp[u] = g;
        }
        rk += __popc(bk[r]);
        rg += __popc(bg[r]);
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        const uint64_t U = (uint32_t)pfx + tk, G = (uint32_t)(pfx >> 32) + tg;
        d_counts[0] = U;
        d_counts[1] = G;
        key_grp[U] = (uint32_t)G;
        grp_start[G] = (uint32_t)n;
    }
}

// ---------------------------------------------------------------------------------------------
// Key sort of the dense path without the library: keys are distinct and their final order is numeric, so nothing has to
// be stable.  Level 1 (top <= 8 bits) is fused into the rank kernel; dense_partition_kernel is level 2 (next <= 8 bits);
// dense_bucket_kernel sorts a final bucket (<= 4096 keys) in shared memory -- the bin variant of index_build.cu on whole
// keys: one counting pass over the next 12 bits with shared-memory atomics, odd-even rounds until nothing moves -- and
// writes the bucket's part of the postings and the CSR arrays (decoupled look-back over the buckets for the key / group
// base, as bucket_finish does).
// ---------------------------------------------------------------------------------------------
// 5 CTAs per SM (48 registers): 0.786 ms on C2 against 0.815 ms at 6 (40 registers, spills) and 0.841 ms at 4 (64 registers)
#ifndef KS_DP_CTAS
#define KS_DP_CTAS 5
#endif
__global__ void __launch_bounds__(DS_THREADS, KS_DP_CTAS)
dense_partition_kernel(const uint64_t* __restrict__ region1, uint32_t cap1, const uint2* __restrict__ chunk_map, DenseScatter sc) {
    __shared__ DenseScatterSmem s_sc;
    __shared__ uint64_t s_dst[DS_TILE];
    const uint2 e = chunk_map[blockIdx.x];  // (dense_chunks_kernel) region | keys << 16, offset in the region
    const uint32_t nv = e.x >> 16, b1 = e.x & 0xffffu;
    if (nv == 0) return;
    const uint64_t* src = region1 + (uint64_t)b1 * cap1 + e.y;
    uint64_t key[DS_ITEMS];
    uint32_t valid = 0;
#pragma unroll
    for (int it = 0; it < DS_ITEMS; it++) {
        const uint32_t i = it * DS_THREADS + threadIdx.x;
        key[it] = 0;
        if (i < nv) { key[it] = src[i]; valid |= 1u << it; }
    }
    scatter_keys(key, valid, sc, b1 << sc.bits, s_sc, s_dst);
}

// CTA shape of the bucket kernel: 512 threads x 4 CTAs per SM (8 keys per thread) or 1024 x 2 (4 keys per thread);
// either way 2048 threads and 32 registers per thread -- with 4 keys per thread the key registers no longer spill
#ifndef KS_DB_THREADS
#define KS_DB_THREADS 512
#endif
constexpr int DB_THREADS = KS_DB_THREADS;
constexpr int DB_CTAS = 2048 / DB_THREADS;
constexpr int DB_WARPS = DB_THREADS / 32;
constexpr int DB_CAP = 4096;
constexpr int DB_ROWS = DB_CAP / DB_THREADS;  // keys per thread
constexpr int DB_BIN_BITS = 13;
constexpr int DB_BINS = 1 << DB_BIN_BITS;  // 16-bit counters, two per word: a bucket holds at most 4096 keys
constexpr int DB_WORDS_PER_THREAD = DB_BINS / 2 / DB_THREADS;  // counter words a thread owns in the scan (8 or 4)
#ifdef KS_DB_NOSLICE
constexpr int DB_HASH_SLICE = 0;
#else
constexpr int DB_HASH_SLICE = 512;         // pattern hashes of a bucket's rank range kept in shared memory when they fit
#endif
constexpr size_t DB_SMEM = (size_t)DB_CAP * 8 + (size_t)DB_BINS * 2 + (size_t)DB_HASH_SLICE * 8;
constexpr uint32_t DB_BIN_SORT_MAX = 32;   // a bin of more keys sends the bucket to the odd-even rounds
constexpr int DB_SLOT_BITS = 12;           // low bits of an item that carry its slot (items are keys shifted up by >= 12)
static_assert(DB_WORDS_PER_THREAD % 4 == 0 && DB_ROWS * DB_WARPS == 128, "scan layout");

struct DenseBucketArgs {
    const uint64_t* region2;   // final buckets, DB_CAP keys each
    const uint32_t* cursor2;   // keys per bucket
    const uint32_t* bstart;    // tuple offset of every bucket
    uint32_t nb;
    int bucket_bits;           // log2(nb): the buckets are the top bucket_bits (<= DENSE_PREFIX_BITS) bits of the hash
    int rem_bits;              // key bits below the bucket bits (<= 64 - DB_SLOT_BITS)
    int loc_bits, pos_bits;
    int rb, parity;            // code bits below the hash prefix; its lowest bit is the pattern / exception parity (DenseSketchArgs)
    const uint64_t* sorted_hash;
    const uint32_t* group_base;
    uint64_t* loc;
    uint64_t* keys;            // segmented layout (index_build.cuh): bucket b's keys / groups at bstart[b] + b + i
    uint32_t *key_grp, *grp_start, *t_size;
    uint64_t* counts;          // [nb] keys | groups << 32 of every bucket
    unsigned long long* d_counts;  // [0] unique keys, [1] groups: accumulated (zeroed per build)
    uint32_t* dir;             // every bucket owns 2^dir_sub + 1 entries
    int dir_sub, dir_shift;
    // exception keys (even codes): their hash is recomputed from the residues
    const uint8_t* residues;
    const uint64_t* offsets;
    int packed;
    uint32_t k;
    const uint32_t* exc_flag;  // device: != 0 when the rank kernel emitted exception keys (picks the instantiation that runs)
    const uint32_t* skip_flag; // device: != 0 when the build is void (unhandled exception / table unusable): nothing to do
};

// One CTA per final bucket (<= 4096 keys, all distinct, final order = numeric order); buckets are independent:
//   1. counting pass over the item's top 13 bits (shared-memory atomics on packed 16-bit counters), scan, scatter: the
//      bucket is then ordered at bin granularity -- a bin is (code, 1/32 of the proteins) on C2 and holds 0-3 keys.  The
//      slot a key drew in its bin rides in the item's low bits (below every key bit), not in a register;
//   2. odd-even transposition rounds until nothing moves (2-3 rounds; measured against counting the smaller keys of the
//      own bin, KS_DB_RANKSORT: 2.79 ms vs 2.99 ms on C2);
//   3. heads from the code / protein fields, postings out in final order;
//   4. keys / key_grp / grp_start in the SEGMENTED layout (index_build.cuh): positions bstart[b] + b + i, known without
//      any other bucket -- no look-back chain, no ticket (round 1 chained the buckets' totals: 19-23 % of the stall
//      samples sat behind it) -- and the bucket's own directory entries.  The pattern hashes of the bucket's prefix range
//      are consecutive entries of sorted_hash, preloaded with one coalesced read at the top.
// EXC: the batch holds exception keys.  Both instantiations are launched back to back; the device flag decides which one
// does the work (the other's CTAs exit at once), so that the host never waits for the rank kernel's flags.  The rare one
// runs a grid-stride loop over the buckets so that its idle launch costs a few hundred CTAs.
template <bool EXC>
__global__ void __launch_bounds__(DB_THREADS, DB_CTAS)
dense_bucket_kernel(DenseBucketArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* B = reinterpret_cast<uint64_t*>(smem_raw);      // [DB_CAP] items: the key's bits below the bucket bits, left-aligned
    uint32_t* cnt = reinterpret_cast<uint32_t*>(B + DB_CAP);  // [DB_BINS / 2] words of two 16-bit bin counters, then offsets
    uint64_t* s_hash = reinterpret_cast<uint64_t*>(cnt + DB_BINS / 2);  // [DB_HASH_SLICE]
    uint64_t* s_keyh = reinterpret_cast<uint64_t*>(cnt);      // [DB_BINS / 4] hashes of the bucket's keys (after the sort)
    constexpr uint32_t KEYH_CAP = DB_BINS / 4;
    const uint16_t* off16 = reinterpret_cast<const uint16_t*>(cnt);
    __shared__ uint32_t s_wsum[DB_WARPS];
    __shared__ uint32_t s_cw[DB_ROWS * DB_WARPS];
    __shared__ uint32_t s_tk, s_tg;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    if (*a.skip_flag != 0) return;  // the build is void (the host takes the general path)
    if ((*a.exc_flag != 0) != EXC) return;  // (uniform over the grid) the other instantiation does the work
    const int up = 64 - a.rem_bits;  // item = key << up: the bucket bits fall off the top
    const int rrb = a.rem_bits - a.loc_bits;  // code bits below the bucket bits (the lowest is the parity)
    // bin = the item's top 13 bits with the parity bit of the code squeezed out when it lies among them (it is 1 for every
    // pattern key and would leave half of the bins empty).  An exception key (parity 0) can then land in a later bin than a
    // larger key: the odd-even rounds run until the whole bucket is in order.
    const int pb = up + a.loc_bits;  // bit of the item that holds the parity of the code
    const bool squeeze = a.parity && pb >= 64 - DB_BIN_BITS && pb < 63;
    const int n_low = squeeze ? DB_BIN_BITS - (63 - pb) : 0;  // bin bits taken from below the parity bit
    auto bin_of = [&](uint64_t it) -> uint32_t {
        if (!squeeze) return (uint32_t)(it >> (64 - DB_BIN_BITS));
        return (uint32_t)(((it >> (pb + 1)) << n_low) | ((it >> (pb - n_low)) & ((1ull << n_low) - 1ull)));
    };
    const int rank_sh = up + a.loc_bits, grp_sh = up + a.pos_bits;  // item >> rank_sh: code bits below the bucket bits
    const uint64_t pos_mask = (1ull << a.pos_bits) - 1ull, pid_mask = (1ull << (a.loc_bits - a.pos_bits)) - 1ull;
    constexpr uint64_t SLOT_MASK = (1ull << DB_SLOT_BITS) - 1ull;
    const uint32_t dir_n = 1u << a.dir_sub;  // directory entries of a bucket (+ its end sentinel)
    auto dir_local = [&](uint64_t h) -> int32_t { return (int32_t)((uint32_t)(h >> a.dir_shift) & (dir_n - 1u)); };
    for (uint32_t b = blockIdx.x; b < a.nb; b += gridDim.x) {
    __syncthreads();  // (grid-stride instantiation) the previous bucket's shared memory is no longer read
    for (uint32_t i = tid; i < DB_BINS / 8; i += DB_THREADS) reinterpret_cast<uint4*>(cnt)[i] = make_uint4(0, 0, 0, 0);
    const uint64_t* src = a.region2 + (uint64_t)b * DB_CAP;
    uint64_t item[DB_ROWS];
    // the first half of the rows is read before the bucket's size is known (a bucket region is DB_CAP keys of allocated
    // memory; what lies past the size is not used): the size and the keys come back in one round trip instead of two
    constexpr int SPEC_ROWS = DB_ROWS / 2;  // (three quarters, and the flag tests after the loads: 2.65-2.67 ms against 2.64 ms)
#pragma unroll
    for (int r = 0; r < SPEC_ROWS; r++) item[r] = src[r * DB_THREADS + tid];
    const uint32_t m = min(a.cursor2[b], (uint32_t)DB_CAP);
    const uint32_t s0 = a.bstart[b];
    const uint32_t kb = s0 + b;  // first key / group position of this bucket's segment
    uint32_t* dir_b = a.dir + (uint64_t)b * (dir_n + 1u);
    // the pattern hashes of this bucket: the prefix groups [b << (16 - bucket_bits), (b + 1) << (16 - bucket_bits))
    const uint32_t pshift = DENSE_PREFIX_BITS - a.bucket_bits;
    const uint32_t slice_base = a.group_base[b << pshift];
    const uint32_t slice_n = a.group_base[(b + 1) << pshift] - slice_base;
    const bool hash_slice = DB_HASH_SLICE > 0 && slice_n <= (uint32_t)DB_HASH_SLICE;
    if (hash_slice) for (uint32_t i = tid; i < slice_n; i += DB_THREADS) s_hash[i] = a.sorted_hash[slice_base + i];
    if (m == 0) {  // an empty key / group segment (the two sentinels), directory entries that all point at it
        if (tid == 0) { a.key_grp[kb] = kb; a.grp_start[kb] = s0; a.counts[b] = 0; }
        for (uint32_t x = tid; x <= dir_n; x += DB_THREADS) dir_b[x] = kb;
        continue;
    }
#pragma unroll
    for (int r = SPEC_ROWS; r < DB_ROWS; r++) {
        if (r * DB_THREADS >= m) break;
        const uint32_t j = r * DB_THREADS + tid;
        if (j < m) item[r] = src[j];
    }
    __syncthreads();  // the counters are zero
#pragma unroll
    for (int r = 0; r < DB_ROWS; r++) {
        const uint32_t j = r * DB_THREADS + tid;
        if (r * DB_THREADS >= m) break;
        if (j < m) {
            item[r] <<= up;
            const uint32_t bin = bin_of(item[r]), sh16 = 16 * (bin & 1u);
            item[r] |= (atomicAdd(&cnt[bin >> 1], 1u << sh16) >> sh16) & 0xffffu;  // no carry: counts <= 4096 = 2^DB_SLOT_BITS
        }
    }
    __syncthreads();
    bool big_bin = false;
    {   // exclusive scan of the 8192 counters; thread t owns DB_WORDS_PER_THREAD consecutive words
        uint32_t w[DB_WORDS_PER_THREAD];
#pragma unroll
        for (int q = 0; q < DB_WORDS_PER_THREAD / 4; q++) {
            const uint4 c = reinterpret_cast<uint4*>(cnt)[tid * (DB_WORDS_PER_THREAD / 4) + q];
            w[4 * q] = c.x; w[4 * q + 1] = c.y; w[4 * q + 2] = c.z; w[4 * q + 3] = c.w;
        }
        uint32_t total = 0;
#pragma unroll
        for (int i = 0; i < DB_WORDS_PER_THREAD; i++) {
            const uint32_t lo = w[i] & 0xffffu, hi = w[i] >> 16;
            total += lo + hi;
            big_bin |= lo > DB_BIN_SORT_MAX || hi > DB_BIN_SORT_MAX;
        }
        uint32_t incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t off = 0;
#pragma unroll
        for (int w2 = 0; w2 < DB_WARPS; w2++) off += w2 < (int)warp ? s_wsum[w2] : 0u;
        uint32_t run = off + incl - total;  // < 4096: the exclusive offsets fit 16 bits as well
#pragma unroll
        for (int i = 0; i < DB_WORDS_PER_THREAD; i++) {
            const uint32_t lo = w[i] & 0xffffu, hi = w[i] >> 16;
            w[i] = run | ((run + lo) << 16);
            run += lo + hi;
        }
#pragma unroll
        for (int q = 0; q < DB_WORDS_PER_THREAD / 4; q++)
            reinterpret_cast<uint4*>(cnt)[tid * (DB_WORDS_PER_THREAD / 4) + q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < DB_ROWS; r++) {
        const uint32_t j = r * DB_THREADS + tid;
        if (r * DB_THREADS >= m) break;
        if (j < m) B[off16[bin_of(item[r])] + (uint32_t)(item[r] & SLOT_MASK)] = item[r];
    }
#ifdef KS_DB_RANKSORT
    const int use_rounds = __syncthreads_or((EXC || big_bin) ? 1 : 0);
#else
    (void)big_bin;
    const int use_rounds = __syncthreads_or(1);
#endif
    if (!use_rounds) {
        // every key finds its place inside its bin by counting the bin's smaller keys (keys are distinct above the slot
        // bits; a bin holds a handful)
#pragma unroll
        for (int r = 0; r < DB_ROWS; r++) {
            const uint32_t j = r * DB_THREADS + tid;
            if (r * DB_THREADS >= m) break;
            if (j < m) {
                const uint32_t bin = bin_of(item[r]);
                const uint32_t lo = off16[bin], hi = bin == DB_BINS - 1 ? m : off16[bin + 1];
                uint32_t at = lo;
                for (uint32_t t = lo; t < hi; t++) at += B[t] < item[r] ? 1u : 0u;
                item[r] = (item[r] & ~SLOT_MASK) | at;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < DB_ROWS; r++) {
            const uint32_t j = r * DB_THREADS + tid;
            if (r * DB_THREADS >= m) break;
            if (j < m) B[(uint32_t)(item[r] & SLOT_MASK)] = item[r];
        }
        __syncthreads();
    } else {
        // odd-even transposition until nothing moves
        const uint32_t np0 = m >> 1, np1 = (m - 1) >> 1;
        ulonglong2* B2 = reinterpret_cast<ulonglong2*>(B);
        int again;
        do {
            int sw = 0;
            for (uint32_t i = tid; i < np0; i += DB_THREADS) {
                const ulonglong2 v = B2[i];
                if (v.x > v.y) { B2[i] = make_ulonglong2(v.y, v.x); sw = 1; }
            }
            __syncthreads();
            for (uint32_t i = tid; i < np1; i += DB_THREADS) {
                const uint64_t x = B[2 * i + 1], y = B[2 * i + 2];
                if (x > y) { B[2 * i + 1] = y; B[2 * i + 2] = x; sw = 1; }
            }
            again = __syncthreads_or(sw);
        } while (again);
    }
    // Exception keys (even code: a window with a residue of neither class, hashed from its bytes, placed between two
    // patterns of its prefix group).  Two different exception hashes can fall between the same two patterns and then share a
    // code: the run of that code is put in (hash, protein, position) order by the thread that owns its first key, and the
    // head tests below compare recomputed hashes.  (rrb >= rb >= 2: the parity bit is always below the bucket bits.)
    auto is_exc = [&](uint64_t it) -> bool { return ((it >> rank_sh) & 1ull) == 0ull; };
    auto hash_of = [&](uint64_t it) -> uint64_t {
        const uint64_t low = it >> up;
        const uint32_t pid = (uint32_t)((low >> a.pos_bits) & pid_mask);
        return dense_window_hash(a.residues, a.packed, a.offsets[pid] + (low & pos_mask), a.k);
    };
    if (EXC) {
        for (uint32_t j = tid; j < m; j += DB_THREADS) {
            const uint64_t it = B[j];
            if (!is_exc(it) || (j && (B[j - 1] >> rank_sh) == (it >> rank_sh))) continue;  // not the first key of a run
            uint32_t e = j + 1;
            while (e < m && (B[e] >> rank_sh) == (it >> rank_sh)) e++;
            if (e - j < 2) continue;
            const uint64_t h0 = hash_of(it);
            bool mixed = false;
            for (uint32_t t = j + 1; t < e && !mixed; t++) mixed = hash_of(B[t]) != h0;
            if (!mixed) continue;
            for (uint32_t t = j + 1; t < e; t++) {  // insertion by (hash, key); hashes recomputed: runs are a handful of keys
                const uint64_t x = B[t], hx = hash_of(x);
                uint32_t u = t;
                while (u > j) {
                    const uint64_t y = B[u - 1], hy = hash_of(y);
                    if (hy < hx || (hy == hx && y < x)) break;
                    B[u] = y;
                    u--;
                }
                B[u] = x;
            }
        }
        __syncthreads();
    }
    uint32_t flags = 0;
#pragma unroll
    for (int r = 0; r < DB_ROWS; r++) {
        const uint32_t j = r * DB_THREADS + tid;
        if (r * DB_THREADS < m) {  // uniform
            bool hk = false, hg = false;
            if (j < m) {
                const uint64_t it = B[j];
                const uint64_t pv = j ? B[j - 1] : ~it;
                hk = j == 0 || (it >> rank_sh) != (pv >> rank_sh);
                if (EXC && !hk && is_exc(it)) hk = hash_of(it) != hash_of(pv);  // same code, maybe another hash
                hg = hk || (it >> grp_sh) != (pv >> grp_sh);
                const uint64_t low = it >> up;  // the key's bits below the bucket bits: code low bits | protein | position
                const uint32_t pid = (uint32_t)((low >> a.pos_bits) & pid_mask);
                a.loc[s0 + j] = ((uint64_t)pid << 32) | (low & pos_mask);
                if (!hg) atomicSub(&a.t_size[pid], 1u);  // the hash again in the same protein: one distinct hash fewer
            }
            flags |= ((hk ? 1u : 0u) | (hg ? 2u : 0u)) << (2 * r);
            const uint32_t ck = __popc(__ballot_sync(0xffffffffu, hk)), cg = __popc(__ballot_sync(0xffffffffu, hg));
            if (lane == 0) s_cw[r * DB_WARPS + warp] = ck | (cg << 16);
        } else if (lane == 0) {
            s_cw[r * DB_WARPS + warp] = 0;
        }
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t v[4], local = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { v[i] = s_cw[lane * 4 + i]; local += v[i]; }
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        uint32_t run = incl - local;
#pragma unroll
        for (int i = 0; i < 4; i++) { s_cw[lane * 4 + i] = run; run += v[i]; }
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {
            const uint32_t tk = tot & 0xffffu, tg = tot >> 16;
            s_tk = tk; s_tg = tg;
            a.counts[b] = (uint64_t)tk | ((uint64_t)tg << 32);
            atomicAdd(a.d_counts, (unsigned long long)tk);
            atomicAdd(a.d_counts + 1, (unsigned long long)tg);
        }
    }
    __syncthreads();
    const uint32_t tk = s_tk, tg = s_tg;
    const uint32_t code_top = b << rrb;  // the code bits that are the bucket index
#pragma unroll
    for (int r = 0; r < DB_ROWS; r++) {
        if (r * DB_THREADS < m) {  // uniform
            const uint32_t j = r * DB_THREADS + tid;
            const bool hk = (flags >> (2 * r)) & 1u, hg = (flags >> (2 * r)) & 2u;
            const uint32_t bk = __ballot_sync(0xffffffffu, hk), bg = __ballot_sync(0xffffffffu, hg);
            const uint32_t pre = s_cw[r * DB_WARPS + warp];
            const uint32_t g = kb + (pre >> 16) + __popc(bg & lt);
            if (hg) a.grp_start[g] = s0 + j;
            if (hk) {
                const uint32_t ul = (pre & 0xffffu) + __popc(bk & lt);
                const uint64_t it = B[j];
                uint64_t h;
                if (EXC && is_exc(it)) {
                    h = hash_of(it);
                } else {
                    const uint32_t code = code_top | (uint32_t)(it >> rank_sh);
                    // (a bucket that is one prefix group -- 2^16 buckets -- knows its group's base already)
                    const uint32_t gb = pshift == 0 ? slice_base : a.group_base[code >> a.rb];
                    const uint32_t idx = gb + ((code & ((1u << a.rb) - 1u)) >> a.parity);
                    h = hash_slice ? s_hash[idx - slice_base] : a.sorted_hash[idx];
                }
                a.keys[kb + ul] = h;
                a.key_grp[kb + ul] = g;
                if (tk <= KEYH_CAP) s_keyh[ul] = h;  // (the bin offsets are no longer needed)
            }
            if (j == m - 1) {  // the segment's sentinels
                a.key_grp[kb + tk] = kb + tg;
                a.grp_start[kb + tg] = s0 + m;
            }
        }
    }
    __syncthreads();
    // the bucket's directory entries: key i owns the entries (local entry of key i - 1, local entry of key i]
    for (uint32_t i = tid; i < tk; i += DB_THREADS) {
        const uint64_t h = tk <= KEYH_CAP ? s_keyh[i] : a.keys[kb + i];
        const int32_t x1 = dir_local(h);
        int32_t x0 = -1;
        if (i) x0 = dir_local(tk <= KEYH_CAP ? s_keyh[i - 1] : a.keys[kb + i - 1]);
        for (int32_t x = x0 + 1; x <= x1; x++) dir_b[x] = kb + i;
        if (i == tk - 1) for (int32_t x = x1 + 1; x <= (int32_t)dir_n; x++) dir_b[x] = kb + tk;  // after the last key, the end
    }
    }  // buckets
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t table_sort_bytes(uint32_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int64_t)n, 0, 64);
    return align256(b);
}

size_t key_sort_bytes(uint64_t n) {
    size_t b = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, b, k, (int64_t)n, 0, 64);
    return align256(b);
}

size_t scan_bytes(uint64_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int64_t)n);
    return align256(b);
}

}  // namespace

size_t dense_table_temp_bytes(uint32_t k) {
    const size_t n = (size_t)1 << k;
    return align256(n * 8) + 2 * align256(n * 4) + table_sort_bytes((uint32_t)n) + 256;
}

#define KS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// Step 1: hash all 2^k patterns, sort the hashes, prefix-group bases, checks.  d_flags (u32[4], [0] and [3] zeroed here):
// [0] != 0 when two patterns share a hash or one hashes to 0, [3] = largest rank of a pattern inside its prefix group
// (the caller reads both and derives the code width `rb` for step 2).  Enqueued on `stream`; no synchronisation.
cudaError_t dense_build_tables(uint32_t k, uint64_t* sorted_hash, uint32_t* group_base, uint32_t* pattern_of_rank, void* temp,
                               size_t temp_bytes, uint32_t* d_flags, cudaStream_t stream, uint64_t* n_launches) {
    const uint32_t n = 1u << k;
    if (temp_bytes < dense_table_temp_bytes(k)) return cudaErrorInvalidValue;
    char* p = (char*)temp;
    uint64_t* hash = (uint64_t*)p; p += align256((size_t)n * 8);
    uint32_t* code = (uint32_t*)p; p += align256((size_t)n * 4);
    size_t sort_bytes = table_sort_bytes(n);
    KS_TRY(cudaMemsetAsync(d_flags, 0, 4, stream));
    KS_TRY(cudaMemsetAsync(d_flags + 3, 0, 4, stream));
    dense_hash_codes_kernel<<<(n + 255) / 256, 256, 0, stream>>>(k, hash, code);
    KS_TRY(cudaGetLastError());
    KS_TRY(cub::DeviceRadixSort::SortPairs(p, sort_bytes, hash, sorted_hash, code, pattern_of_rank, (int64_t)n, 0, 64, stream));
    dense_group_base_kernel<<<((1u << DENSE_PREFIX_BITS) + 1 + 255) / 256, 256, 0, stream>>>(n, sorted_hash, group_base);
    dense_check_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sorted_hash, group_base, d_flags);
    if (n_launches) *n_launches += 3 + 10;
    return cudaGetLastError();
}

// Step 2: the pattern -> code table for the code width and layout the host derived (again whenever the layout changes).
cudaError_t dense_build_codes(uint32_t k, const uint64_t* sorted_hash, const uint32_t* group_base, const uint32_t* pattern_of_rank,
                              int rb, int parity, uint32_t* code_of_pattern, cudaStream_t stream, uint64_t* n_launches) {
    const uint32_t n = 1u << k;
    dense_code_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sorted_hash, pattern_of_rank, group_base, rb, parity, code_of_pattern);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

DenseSortPlan dense_sort_plan(uint64_t n, int rank_bits, int key_bits, int k) {
    DenseSortPlan p;
    int total = 0;
    while (total <= 2 * DS_MAX_BITS && (n >> total) > 3072) total++;
    // a bucket is a hash-prefix range: it holds Poisson(2^k / 2^total) patterns of ~n / 2^k keys each.  With few patterns per
    // bucket the loads are uneven: take more buckets until mean + 6 sigma fits one (or the prefix bits run out: a bucket
    // that overflows sends the batch to the general path)
    // ... and with so few patterns that one of them fills a bucket on its own (hp k <= 13 at 10 M residues) the library
    // sorts the keys on their code bits (stable; rows of one pattern are already in (protein, position) order)
    if ((double)n / std::ldexp(1.0, k) > 1024.0) return p;
    while (total < DENSE_PREFIX_BITS) {
        const double lambda = std::ldexp(1.0, k - total), per_pattern = (double)n / std::ldexp(1.0, k);
        if (lambda * per_pattern + 6.0 * per_pattern * std::sqrt(lambda > 1.0 ? lambda : 1.0) <= (double)DB_CAP) break;
        total++;
    }
    if (total > 2 * DS_MAX_BITS || total > DENSE_PREFIX_BITS || total > rank_bits || n == 0) return p;  // custom == 0
    if (key_bits - total > 64 - DB_SLOT_BITS) return p;  // the bucket kernel keeps a key's slot below its bits
    p.custom = 1;
    p.total = total;
    p.l1 = total < DS_MAX_BITS ? total : DS_MAX_BITS;
    p.l2 = total - p.l1;
    p.cap1 = p.l2 ? (uint32_t)((n >> p.l1) + (n >> (p.l1 + 3)) + 16384) : (uint32_t)DB_CAP;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes); return o; };
    p.off_region1 = take(((size_t)p.cap1 << p.l1) * 8);
    p.off_region2 = p.l2 ? take(((size_t)DB_CAP << total) * 8) : p.off_region1;
    p.off_small = off;
    p.off_cursor1 = take(((size_t)1 << p.l1) * 4);
    p.off_cursor2 = p.l2 ? take(((size_t)1 << total) * 4) : p.off_cursor1;
    p.off_overflow = take(8);
    p.small_bytes = off - p.off_small;  // everything above is zeroed before a build
    p.max_chunks = (uint32_t)(n / DS_TILE + ((size_t)1 << p.l1) + 1);  // every region's last chunk may be partial
    p.off_chunks = take((size_t)p.max_chunks * 8);
    p.off_bstart = take((((size_t)1 << total) + 1) * 4);
    p.off_counts = take(((size_t)1 << total) * 8);
    p.bytes = off;
    return p;
}

size_t dense_csr_temp_bytes(uint64_t n) {
    const uint64_t nt = (n + DC_TILE - 1) / DC_TILE;
    return key_sort_bytes(n) + 2 * align256((nt + 1) * 8) + scan_bytes(nt + 1) + 256;
}

cudaError_t dense_build_csr(const DenseCsrArgs& a, cudaStream_t stream, uint64_t* sort_launches, uint64_t* csr_launches) {
    if (a.out_dir_sub) *a.out_dir_sub = DIR_SUB_COMPACT;
    if (a.out_seg_nb) *a.out_seg_nb = 0;
    const uint64_t n = a.n;
    const int loc_bits = a.pid_bits + a.pos_bits;
    if (a.n_prot) {
        dense_windows_kernel<<<(a.n_prot + 255) / 256, 256, 0, stream>>>(a.offsets, a.n_prot, a.k, a.t_abund, a.t_size);
        KS_TRY(cudaGetLastError());
        if (csr_launches) *csr_launches += 1;
    }
    if (n == 0) {
        KS_TRY(cudaMemsetAsync(a.d_counts, 0, 16, stream));
        KS_TRY(cudaMemsetAsync(a.key_grp, 0, 4, stream));
        KS_TRY(cudaMemsetAsync(a.grp_start, 0, 4, stream));
        if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
        return cudaSuccess;
    }
    if (a.plan.custom) {
        const DenseSortPlan& pl = a.plan;
        char* w = (char*)a.work;
        uint64_t* region1 = (uint64_t*)(w + pl.off_region1);
        uint64_t* region2 = (uint64_t*)(w + pl.off_region2);
        uint32_t* cursor1 = (uint32_t*)(w + pl.off_cursor1);
        uint32_t* cursor2 = (uint32_t*)(w + pl.off_cursor2);
        uint2* chunk_map = (uint2*)(w + pl.off_chunks);
        uint32_t* bstart = (uint32_t*)(w + pl.off_bstart);
        const uint32_t nb = 1u << pl.total;
        const int total_bits = a.rank_bits + loc_bits;
        if (pl.l2) {  // second level: every first-level region into 2^l2 final buckets
            dense_chunks_kernel<<<1, 256, 0, stream>>>(cursor1, 1u << pl.l1, pl.cap1, chunk_map, pl.max_chunks);
            DenseScatter sc;
            sc.out = region2; sc.cursor = cursor2; sc.cap = (uint32_t)DB_CAP; sc.shift = total_bits - pl.total; sc.bits = pl.l2;
            sc.overflow = a.overflow ? a.overflow : (uint32_t*)(w + pl.off_overflow);
            dense_partition_kernel<<<pl.max_chunks, DS_THREADS, 0, stream>>>(region1, pl.cap1, chunk_map, sc);
            KS_TRY(cudaGetLastError());
            if (sort_launches) *sort_launches += 2;
        }
        dense_bucket_offsets_kernel<<<1, 1024, 0, stream>>>(cursor2, nb, (uint32_t)DB_CAP, bstart);
        if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
        if (a.dir_bits < pl.total) return cudaErrorInvalidValue;  // (the caller sizes the directory from the plan)
        if (!a.counts_zeroed) KS_TRY(cudaMemsetAsync(a.d_counts, 0, 16, stream));
        DenseBucketArgs ba;
        ba.region2 = region2; ba.cursor2 = cursor2; ba.bstart = bstart; ba.nb = nb; ba.bucket_bits = pl.total;
        ba.rem_bits = total_bits - pl.total; ba.loc_bits = loc_bits; ba.pos_bits = a.pos_bits; ba.rb = a.rb; ba.parity = a.parity;
        ba.sorted_hash = a.sorted_hash; ba.group_base = a.group_base;
        ba.loc = a.loc; ba.keys = a.keys; ba.key_grp = a.key_grp; ba.grp_start = a.grp_start; ba.t_size = a.t_size;
        ba.counts = (uint64_t*)(w + pl.off_counts); ba.d_counts = (unsigned long long*)a.d_counts;
        ba.dir = a.dir; ba.dir_sub = a.dir_bits - pl.total; ba.dir_shift = 64 - a.dir_bits;
        ba.residues = a.residues; ba.offsets = a.offsets; ba.packed = a.packed; ba.k = a.k;
        ba.exc_flag = a.exc_flag; ba.skip_flag = a.skip_flag;
        static thread_local int attr_device = -1;  // (the attribute is per device: once per host thread and device)
        int dev = 0;
        KS_TRY(cudaGetDevice(&dev));
        if (attr_device != dev) {
            KS_TRY(cudaFuncSetAttribute(dense_bucket_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DB_SMEM));
            KS_TRY(cudaFuncSetAttribute(dense_bucket_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DB_SMEM));
            attr_device = dev;
        }
        // both instantiations, back to back: the rank kernel's device flag picks the one that works -- the host does not
        // wait for the flag (the exception instantiation strides over the buckets: idle, it is a few hundred CTAs that exit)
        dense_bucket_kernel<false><<<nb, DB_THREADS, DB_SMEM, stream>>>(ba);
        dense_bucket_kernel<true><<<std::min<unsigned>(nb, 148u * DB_CTAS), DB_THREADS, DB_SMEM, stream>>>(ba);
        if (sort_launches) *sort_launches += 3;
        if (a.out_dir_sub) *a.out_dir_sub = ba.dir_sub;
        if (a.out_seg_nb) *a.out_seg_nb = nb;
        if (a.out_seg_start) *a.out_seg_start = bstart;
        if (a.out_seg_counts) *a.out_seg_counts = ba.counts;
        return cudaGetLastError();
    }
    char* p = (char*)a.temp;
    size_t sort_bytes = key_sort_bytes(n);
    void* sort_temp = p; p += sort_bytes;
    const uint64_t nt = (n + DC_TILE - 1) / DC_TILE;
    uint64_t* tile_counts = (uint64_t*)p; p += align256((nt + 1) * 8);
    uint64_t* tile_prefix = (uint64_t*)p; p += align256((nt + 1) * 8);
    size_t sbytes = scan_bytes(nt + 1);
    void* scan_temp = p;
    // 1. stable library sort on the rank bits: keys arrive in (protein, position) order and leave in (rank, protein, position)
    cub::DoubleBuffer<uint64_t> kb(a.keys_a, a.keys_b);
    KS_TRY(cub::DeviceRadixSort::SortKeys(sort_temp, sort_bytes, kb, (int64_t)n, loc_bits, loc_bits + a.rank_bits, stream));
    if (sort_launches) *sort_launches += 2 + (a.rank_bits + 7) / 8;
    if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
    const uint64_t* sorted = kb.Current();
    // 2. heads per tile, scan, CSR write
    dense_count_kernel<<<(unsigned)nt, DC_THREADS, 0, stream>>>(sorted, n, loc_bits, a.pos_bits, tile_counts);
    KS_TRY(cudaGetLastError());
    KS_TRY(cub::DeviceScan::ExclusiveSum(scan_temp, sbytes, tile_counts, tile_prefix, (int64_t)(nt + 1), stream));
    dense_write_kernel<<<(unsigned)nt, DC_THREADS, 0, stream>>>(sorted, n, loc_bits, a.pos_bits, tile_prefix, a.sorted_hash, a.group_base,
                                                                a.rb, a.parity, a.loc, a.keys, a.key_grp, a.grp_start, a.t_size, a.d_counts);
    if (csr_launches) *csr_launches += 4;
    if (a.out_dir_sub) *a.out_dir_sub = DIR_SUB_COMPACT;  // compact layout: the caller builds the directory (launch_directory)
    if (a.out_seg_nb) *a.out_seg_nb = 0;
    return cudaGetLastError();
#undef KS_TRY
}

}  // namespace ks
