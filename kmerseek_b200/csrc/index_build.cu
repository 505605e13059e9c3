// Index build for sm_100a: MSD sort of the sketch tuples by hash and chain-free CSR construction.
//
// Replaces, on the hot path, the reference's per-protein sorted Vec inserts (sourmash KmerMinHash via
// src/rust/signature.rs:273-274), the nested HashMap position records (src/rust/index.rs:770-780) and the
// serial sorted-Vec merge into combined_minhash (src/rust/index.rs:824-827).
//
// Pipeline of the general path (hp with k <= 24 normally takes the dense path of dense.cu):
//   1. partition the tuples by the top `tb` bits of the hash (library onesweep passes over those bits only)
//   2. bucket sort: one CTA sorts one bucket (a few thousand tuples) entirely in shared memory --
//      bucket_sort_bin_kernel when hashes rarely repeat, bucket_sort_rep_kernel when they do -- and both end in
//      bucket_finish: the postings go back in final order, the bucket's unique hashes / (hash, protein) groups are
//      counted, and after a decoupled look-back over the buckets for its key / group base the bucket writes its part
//      of keys / key_grp / grp_start and of the bucket directory straight from shared memory
//   3. only when a bucket is too large for shared memory (heavy repeats of one hash): library sort of those ranges,
//      then scan_counts_kernel + csr_write_kernel + dir_kernel build the CSR in separate passes
// Small inputs take the library sort for all bits and step 3 over fixed 4096-tuple ranges.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "dense_scatter.cuh"
#include "index_build.cuh"

namespace ks {

namespace {

constexpr int LS_THREADS = 512;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_CAP = 4096;  // tuples per bucket that fit the shared-memory layout (12 index bits)

__global__ void protein_abund_kernel(const uint64_t* __restrict__ loc, uint64_t n, uint32_t n_prot,
                                     uint32_t* __restrict__ t_abund, uint32_t* __restrict__ t_size) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    // tuples are ordered by (protein, pos): count = lower_bound((p+1)<<32) - lower_bound(p<<32); a lane takes its
    // upper bound from the lane above (the last lane of a warp searches twice)
    auto lb = [&](uint64_t key) {
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (loc[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const uint64_t mine = p <= n_prot ? lb((uint64_t)p << 32) : n;
    uint64_t next = __shfl_down_sync(0xffffffffu, mine, 1);
    if (lane == 31 && p < n_prot) next = lb(((uint64_t)p + 1) << 32);
    if (p >= n_prot) return;
    const uint32_t c = (uint32_t)(next - mine);
    t_abund[p] = c;
    t_size[p] = c;  // repeats of a (hash, protein) pair are subtracted while the groups are counted
}

__global__ void bucket_start_kernel(const uint64_t* __restrict__ hash, uint64_t n, int lz, int tb,
                                    uint32_t* __restrict__ start, uint32_t* __restrict__ oversize) {
    const uint32_t nb = 1u << tb;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    if (b == 0) oversize[0] = 0;
    if (b == nb) { start[b] = (uint32_t)n; return; }
    uint64_t lo = 0, hi = n;  // first i with bucket(hash[i]) >= b
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint32_t bm = tb ? (uint32_t)((hash[mid] << lz) >> (64 - tb)) : 0u;
        if (bm < b) lo = mid + 1; else hi = mid;
    }
    start[b] = (uint32_t)lo;
}

__global__ void fixed_ranges_kernel(uint64_t n, uint32_t nb, uint32_t* __restrict__ start) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    const uint64_t s = (uint64_t)b * LS_CAP;
    start[b] = (uint32_t)(s < n ? s : n);
}

// Buckets too large for shared memory (only heavy repeats of one hash can do that) are counted for the
// host, which sorts those ranges with the library radix sort.
__global__ void oversize_count_kernel(const uint32_t* __restrict__ start, uint32_t nb, uint32_t* __restrict__ oversize) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    if (start[b + 1] - start[b] > (uint32_t)LS_CAP) atomicAdd(&oversize[0], 1u);
}

// ---------------------------------------------------------------------------------------------
// Shared tail of both bucket-sort kernels: stream the sorted bucket back to HBM, count its unique hashes and
// (hash, protein) groups, and -- when no bucket is oversize -- write the bucket's part of the CSR arrays too, in the
// SEGMENTED layout (index_build.cuh): the bucket's keys and groups live at positions start[b] + b + i, which every
// bucket knows without looking at any other bucket.  There is no chain, no look-back and no ticket between buckets
// (round 1 chained the buckets' key / group totals with a decoupled look-back: 19 % of the kernel's stall samples sat
// behind it, and every slow bucket held up all of its successors).  The totals are two atomics per bucket.
// ---------------------------------------------------------------------------------------------
struct CsrOut {
    const uint32_t* oversize;  // [0] = number of oversize buckets; the CSR is written here only when it is 0
    uint64_t* keys;
    uint32_t* key_grp;
    uint32_t* grp_start;
    unsigned long long* d_counts;  // [0] unique keys, [1] groups: accumulated (zeroed per build)
    uint32_t nb;
    uint32_t* dir;  // every sort bucket owns 2^dir_sub + 1 consecutive entries (the last one is its end sentinel)
    int dir_sub;
};

// dir[x] = position of the first key whose directory bucket is >= x.  Inside a sort bucket (`dir_b`) the key at position u
// whose item has the local entry x1, and whose predecessor has x0, owns the entries (x0, x1].
__device__ __forceinline__ void dir_fill(uint32_t* dir_b, int32_t x0, int32_t x1, uint32_t u) {
    for (int32_t x = x0 + 1; x <= x1; x++) dir_b[x] = u;
}

// An empty bucket: an empty key / group segment (the two sentinels) and directory entries that all point at it.
__device__ __forceinline__ void bucket_empty(uint32_t b, uint32_t s, const CsrOut& f) {
    if (f.oversize[0] != 0) return;
    const uint32_t kb = s + b;
    if (threadIdx.x == 0) { f.key_grp[kb] = kb; f.grp_start[kb] = s; }
    uint32_t* dir_b = f.dir + (uint64_t)b * ((1u << f.dir_sub) + 1u);
    for (uint32_t x = threadIdx.x; x <= (1u << f.dir_sub); x += blockDim.x) dir_b[x] = kb;
}

// The loc gathers of a thread are issued back to back; the protein ids are parked in shared memory (`spid`, 4 bytes per
// tuple of scratch), so a head test reads two neighbouring slots.  No sorted hash column on the fused path: the CSR
// arrays carry the hashes (keys) and nothing on the hot path reads hash[i] of the sorted tuples (expand_sorted_hash
// rebuilds the column for the export calls).
// `n_oversize` = f.oversize[0], read by the caller at the head of the kernel (read here, the first use waited a full
// round trip: 8 % of the bin kernel's stall samples).  DirFirst: the first tuple of the bucket's directory entries when
// the caller knows them from its bin offsets (bucket_sort_bin_kernel: a directory entry is a prefix of a bin index) --
// the directory is then one coalesced write per entry instead of a per-key loop (11 % of the bin kernel's instructions).
struct DirFirst {
    bool from_bins = false;
    uint32_t first[2] = {0, 0};  // first tuple of entries tid and tid + LS_THREADS
};

__device__ __forceinline__ void bucket_finish(const uint64_t* items, uint32_t* spid, uint32_t m, uint32_t s_in, uint32_t s, uint32_t b,
                                                  int lz, int tb, const uint64_t* __restrict__ in_loc,
                                                  uint64_t* __restrict__ out_hash, uint64_t* __restrict__ out_loc,
                                                  uint64_t* __restrict__ counts, uint32_t* __restrict__ t_size,
                                                  const CsrOut& f, uint32_t n_oversize, const DirFirst& df) {
    __shared__ uint32_t s_cw[8 * LS_WARPS];  // per (row, warp): key heads | group heads << 16, then their prefix
    __shared__ uint32_t s_tk, s_tg;          // unique keys / groups of this bucket
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t top = tb ? ((uint64_t)b << (64 - tb)) : 0ull;
    const bool fused = n_oversize == 0;
    uint64_t loc[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t j = r * LS_THREADS + tid;
        if (r * LS_THREADS >= m) break;  // uniform
        if (j < m) loc[r] = in_loc[s_in + (uint32_t)(items[j] & 0xfffu)];  // the bucket's own window of the input (L1/L2)
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t j = r * LS_THREADS + tid;
        if (r * LS_THREADS >= m) break;  // uniform
        if (j < m) spid[j] = (uint32_t)(loc[r] >> 32);
    }
    __syncthreads();
    uint32_t flags = 0;  // 2 bits per row: this thread's element is a key head / a group head
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t j = r * LS_THREADS + tid;
        if (r * LS_THREADS < m) {  // uniform
            bool hk = false, hg = false;
            if (j < m) {
                const uint32_t pid = (uint32_t)(loc[r] >> 32);
                hk = j == 0 || ((items[j - 1] ^ items[j]) >> 12) != 0;
                hg = hk || spid[j - 1] != pid;
                if (!hg) atomicSub(&t_size[pid], 1u);
            }
            flags |= ((hk ? 1u : 0u) | (hg ? 2u : 0u)) << (2 * r);
            const uint32_t ck = __popc(__ballot_sync(0xffffffffu, hk)), cg = __popc(__ballot_sync(0xffffffffu, hg));
            if (lane == 0) s_cw[r * LS_WARPS + warp] = ck | (cg << 16);
        } else if (lane == 0) {
            s_cw[r * LS_WARPS + warp] = 0;
        }
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive prefix over the (row, warp) counts, 4 entries per lane; both halves stay below 2^16
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
        const uint32_t tk = tot & 0xffffu, tg = tot >> 16;
        if (lane == 0) {
            counts[b] = (uint64_t)tk | ((uint64_t)tg << 32);
            s_tk = tk; s_tg = tg;
            if (fused) { atomicAdd(f.d_counts, (unsigned long long)tk); atomicAdd(f.d_counts + 1, (unsigned long long)tg); }
        }
    }
    // the tuples go back to HBM in final order
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t j = r * LS_THREADS + tid;
        if (r * LS_THREADS >= m) break;  // uniform
        if (j < m) {
            out_loc[s + j] = loc[r];
            if (!fused) out_hash[s + j] = (top | ((items[j] & ~0xfffull) >> tb)) >> lz;  // the fallback passes read it
        }
    }
    if (!fused) return;  // an oversize bucket exists: the CSR is written by csr_write_kernel after the fallback
    __syncthreads();
    const uint32_t kb = s + b;  // first key / group position of this bucket's segment
    const uint32_t tk = s_tk, tg = s_tg;
    uint32_t* dir_b = f.dir + (uint64_t)b * ((1u << f.dir_sub) + 1u);
    // (dir_sub <= 24: the entry is a 32-bit shift of the item's high word)
    const uint32_t dir_sh = 32u - (uint32_t)f.dir_sub;
    auto dir_local = [&](uint64_t item) -> uint32_t { return f.dir_sub > 0 ? (uint32_t)(item >> 32) >> dir_sh : 0u; };
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r * LS_THREADS < m) {  // uniform
            const uint32_t j = r * LS_THREADS + tid;
            const bool hk = (flags >> (2 * r)) & 1u, hg = (flags >> (2 * r)) & 2u;
            const uint32_t bk = __ballot_sync(0xffffffffu, hk), bg = __ballot_sync(0xffffffffu, hg);
            const uint32_t pre = s_cw[r * LS_WARPS + warp];
            const uint32_t g = kb + (pre >> 16) + __popc(bg & lt);
            if (hg) f.grp_start[g] = s + j;
            if (hk) {
                const uint32_t u = kb + (pre & 0xffffu) + __popc(bk & lt);
                const uint64_t full = top | ((items[j] & ~0xfffull) >> tb);
                f.keys[u] = full >> lz;
                f.key_grp[u] = g;
                if (df.from_bins) spid[j] = u;  // (the protein ids are no longer needed) key position of a head tuple
                // directory entries this key owns: the bucket's 2^dir_sub entries are indexed by the item's top dir_sub bits
                else dir_fill(dir_b, j ? (int32_t)dir_local(items[j - 1]) : -1, (int32_t)dir_local(items[j]), u);
            }
            if (j == m - 1) {  // the segment's sentinels, and the directory entries after the last key (its own end included)
                f.key_grp[kb + tk] = kb + tg;
                f.grp_start[kb + tg] = s + m;
                if (!df.from_bins) dir_fill(dir_b, (int32_t)dir_local(items[j]), (int32_t)(1u << f.dir_sub), kb + tk);
            }
        }
    }
    if (df.from_bins) {
        // dir[x] = the first key whose entry is >= x = the key of the first tuple of entry x (a head: its predecessor lies
        // in a lower bin), or the segment's end when no tuple is left
        __syncthreads();
        const uint32_t dir_n = 1u << f.dir_sub;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint32_t x = tid + i * LS_THREADS;
            if (x < dir_n) dir_b[x] = df.first[i] < m ? spid[df.first[i]] : kb + tk;
        }
        if (tid == 0) dir_b[dir_n] = kb + tk;
    }
}

// ---------------------------------------------------------------------------------------------
// Bucket-local sort for repeat-heavy inputs (small k-mer space, e.g. hp k=24: a bucket holds a few hundred
// distinct hashes, each about ten times).  Comparison-based clean-up is expensive there (lists that interleave
// two repeated hashes), so this variant resolves 16 more bits with two stable 8-bit counting passes, after
// which runs that still hold an inversion are rare; those are re-ranked by a warp each.
// ---------------------------------------------------------------------------------------------
constexpr size_t LS_SMEM_REP = (size_t)LS_CAP * 8 * 2 + LS_WARPS * 256 * 2 + 256 * 4 + 64;  // 73 KB: 3 CTAs per SM

__global__ void __launch_bounds__(LS_THREADS, 3)
bucket_sort_rep_kernel(const uint64_t* __restrict__ in_hash, const uint64_t* __restrict__ in_loc,
                   uint64_t* __restrict__ out_hash, uint64_t* __restrict__ out_loc,
                   const uint32_t* __restrict__ start, int lz, int tb, uint64_t* __restrict__ counts,
                   uint32_t* __restrict__ t_size, CsrOut f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* A = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* B = A + LS_CAP;
    uint16_t* cnt = reinterpret_cast<uint16_t*>(B + LS_CAP);              // [LS_WARPS][256] warp-private counters
    uint32_t* dbase = reinterpret_cast<uint32_t*>(cnt + LS_WARPS * 256);  // [256] exclusive digit offsets
    __shared__ uint32_t s_wsum[8];
    __shared__ uint32_t s_ninv;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = blockIdx.x;  // buckets are independent: no ticket, no order
    const uint32_t s = start[b], e = start[b + 1];
    const uint32_t m = e - s;
    if (m == 0) { if (tid == 0) counts[b] = 0; bucket_empty(b, s, f); return; }
    if (m > (uint32_t)LS_CAP) return;  // oversize: sorted and counted by the host-driven fallback
    const int sh = lz + tb;  // >= 12 on this path: the low 12 bits of (hash << sh) are zero and carry the index

    // rows of 32 consecutive elements; every warp owns R consecutive rows
    const uint32_t R = (m + LS_THREADS - 1) / LS_THREADS;  // 1..8
    const uint32_t padded = R * LS_THREADS;

    // every warp loads the rows it owns (the same rows it counts below: no block-wide barrier in between)
    for (uint32_t r = 0; r < R; r++) {
        const uint32_t j = (warp * R + r) * 32 + lane;
        uint64_t item = ~0ull;
        if (j < m) item = ((in_hash[s + j] << sh) & ~0xfffull) | j;
        A[j] = item;
    }
    if (tid == 0) s_ninv = 0;

    uint64_t* src = A;
    uint64_t* dst = B;
    const uint32_t lt = (1u << lane) - 1u;
    uint16_t* my = cnt + warp * 256;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const int dshift = 48 + 8 * pass;
        // the digit counters are warp-private: the owning warp clears them, no block-wide barrier needed (the rows
        // read next were written by this warp in pass 0, and before the barrier that ends pass 0 in pass 1)
#pragma unroll
        for (int i = 0; i < 4; i++) reinterpret_cast<uint32_t*>(my)[lane + 32 * i] = 0;
        __syncwarp();
        // sweep 1: groups of equal digits within a row; digit, rank in the group and group size are kept for sweep 2
        uint32_t info[8];
#pragma unroll
        for (uint32_t r = 0; r < 8; r++) {
            info[r] = 0;
            if (r < R) {
                const uint32_t d = (uint32_t)(src[(warp * R + r) * 32 + lane] >> dshift) & 255u;
                // lanes of the row that share the digit: MATCH.ANY (ADU pipe, ~40 cycles per warp) in the first
                // pass, 8 ballots (ALU pipe) in the second, so that neither pipe carries both passes
                uint32_t peers;
                if (true) {
                    peers = __match_any_sync(0xffffffffu, d);
                } else {
                    peers = 0xffffffffu;
#pragma unroll
                    for (int bit = 0; bit < 8; bit++) {
                        const bool on = (d >> bit) & 1u;
                        const uint32_t v = __ballot_sync(0xffffffffu, on);
                        peers &= on ? v : ~v;
                    }
                }
                const uint32_t rank = __popc(peers & lt), size = __popc(peers);
                if (rank == 0) my[d] += (uint16_t)size;
                info[r] = d | (rank << 8) | (size << 16);
                __syncwarp();
            }
        }
        __syncthreads();
        // exclusive scan: over warps within a digit, then over digits
        uint32_t total = 0;
        if (tid < 256) {
#pragma unroll
            for (int w = 0; w < LS_WARPS; w++) {
                uint32_t c = cnt[w * 256 + tid];
                cnt[w * 256 + tid] = (uint16_t)total;
                total += c;
            }
            uint32_t incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            if (lane == 31) s_wsum[warp] = incl;
            total = incl - total;  // exclusive within this warp of digits
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t off = 0;
            for (uint32_t w = 0; w < warp; w++) off += s_wsum[w];
            dbase[tid] = off + total;
        }
        __syncthreads();
        // sweep 2: stable scatter
#pragma unroll
        for (uint32_t r = 0; r < 8; r++) {
            if (r < R) {
                const uint64_t item = src[(warp * R + r) * 32 + lane];
                const uint32_t d = info[r] & 255u, rank = (info[r] >> 8) & 255u, size = info[r] >> 16;
                const uint32_t base = dbase[d] + my[d];
                __syncwarp();
                if (rank == 0) my[d] += (uint16_t)size;
                __syncwarp();
                dst[base + rank] = item;
            }
        }
        __syncthreads();
        uint64_t* t = src; src = dst; dst = t;
    }
    // src == A.  Items are ordered by the 16 digit bits.  Runs that share those bits and hold an inversion are
    // re-ranked by a warp each: every lane counts the run's items below its own (items are distinct).
    uint32_t* headmap = reinterpret_cast<uint32_t*>(B);  // bit j: item j starts a run        [padded / 32 words]
    uint32_t* dirty = headmap + LS_CAP / 32;              // bit j: the run that starts at j has an inversion
    uint32_t* runs = dirty + LS_CAP / 32;                 // run starts to fix                  [<= m]
    for (uint32_t j = tid; j < padded; j += LS_THREADS) {
        bool head = false, inv = false;
        if (j < m) {
            const uint64_t c = src[j];
            if (j == 0) head = true;
            else {
                const uint64_t a = src[j - 1];
                head = ((a ^ c) >> 48) != 0;
                inv = !head && a > c;
            }
        }
        const uint32_t hb = __ballot_sync(0xffffffffu, head);
        const uint32_t ib = __ballot_sync(0xffffffffu, inv);
        if (lane == 0) { headmap[j >> 5] = hb; dirty[j >> 5] = 0; }
        if (ib && lane == 0) atomicAdd(&s_ninv, 1u);
    }
    __syncthreads();
    const uint32_t any_inv = s_ninv;
    __syncthreads();
    if (any_inv) {  // uniform
        if (tid == 0) s_ninv = 0;
        __syncthreads();
        // mark the run of every inversion: its start is the last head bit at or below j
        for (uint32_t j = tid + 1; j < m; j += LS_THREADS) {
            const uint64_t a = src[j - 1], c = src[j];
            if (((a ^ c) >> 48) == 0 && a > c) {
                uint32_t w = j >> 5;
                uint32_t bits = headmap[w] & (0xffffffffu >> (31 - (j & 31)));
                while (bits == 0) bits = headmap[--w];
                const uint32_t st = (w << 5) + 31 - __clz(bits);
                atomicOr(&dirty[st >> 5], 1u << (st & 31));
            }
        }
        __syncthreads();
        for (uint32_t w = tid; w < padded / 32; w += LS_THREADS) {
            uint32_t bits = dirty[w];
            while (bits) {
                const uint32_t bpos = __ffs(bits) - 1;
                bits &= bits - 1;
                runs[atomicAdd(&s_ninv, 1u)] = (w << 5) + bpos;
            }
        }
        __syncthreads();
        const uint32_t nruns = s_ninv;
        for (uint32_t k = warp; k < nruns; k += LS_WARPS) {
            const uint32_t a0 = runs[k];
            // run end: next head bit after a0 (or m)
            uint32_t w = a0 >> 5;
            uint32_t bits = (a0 & 31) == 31 ? 0u : (headmap[w] & (0xffffffffu << ((a0 & 31) + 1)));
            while (bits == 0 && ++w < padded / 32) bits = headmap[w];
            uint32_t b0 = bits ? (w << 5) + __ffs(bits) - 1 : m;
            if (b0 > m) b0 = m;
            const uint32_t r = b0 - a0;
            if (r <= 128) {
                uint64_t mine[4];
                uint32_t rank[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t t = a0 + lane + 32 * q;
                    mine[q] = t < b0 ? src[t] : ~0ull;
                    rank[q] = 0;
                }
                for (uint32_t t = a0; t < b0; t++) {
                    const uint64_t x = src[t];
#pragma unroll
                    for (int q = 0; q < 4; q++) rank[q] += x < mine[q];
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (a0 + lane + 32 * q < b0) src[a0 + rank[q]] = mine[q];
            } else if (lane == 0) {  // long mixed run (pathological input): insertion sort by one thread
                for (uint32_t t = a0 + 1; t < b0; t++) {
                    const uint64_t x = src[t];
                    uint32_t u = t;
                    while (u > a0 && src[u - 1] > x) { src[u] = src[u - 1]; u--; }
                    src[u] = x;
                }
            }
        }
        __syncthreads();
    }

    // src == A: the other item buffer is free and parks the protein ids of the tail
    bucket_finish(src, reinterpret_cast<uint32_t*>(dst), m, s, s, b, lz, tb, in_loc, out_hash, out_loc, counts, t_size, f, f.oversize[0],
                  DirFirst());
}


// ---------------------------------------------------------------------------------------------
// Bucket-local sort, bin variant (the default): ONE unstable counting pass over the next 12 key bits with
// shared-memory atomics, then odd-even transposition rounds on the full items until nothing moves.
//
// The item carries the tuple's index in the bucket in its low 12 bits, so sorting items by value IS the stable
// sort: the counting pass does not have to be stable, which removes the warp-private counters, the MATCH.ANY
// ranking, the (digit, warp) scan and the second pass of the other two variants.  A bin holds the repeats of one
// hash (hp k=24: ~11) or 0-2 unrelated items (uniform hashes).  Tuples are taken in rows of 512 consecutive
// items; with a barrier after every row (STEPPED, repeat-heavy input) a bin is filled row by row, so an item is
// at most a few slots from its final position and a handful of rounds, perfectly balanced over the threads,
// finish the sort.  The same code is exact for any input: a bucket full of one hash just takes more rounds.
// ---------------------------------------------------------------------------------------------
constexpr int BN_BITS = 12;
constexpr int BN_BINS = 1 << BN_BITS;
constexpr int BN_PER_THREAD = BN_BINS / LS_THREADS;  // 8 consecutive bins per thread
constexpr size_t BN_SMEM = (size_t)LS_CAP * 8 + (size_t)BN_BINS * 4;  // 48 KB
static_assert(BN_PER_THREAD == 8, "two 16-byte loads per thread in the scan");

// 3 CTAs per SM: measured against 2 (2.14 ms on the 100 M-residue target run) and 4 (1.85 ms; the loc gather loses its
// L1 lines to the larger shared-memory carve-out) -- 1.73 ms
// SCATTERED: the bucket's tuples come from an unstable partition (dense_scatter.cuh): bucket b holds `cursor[b]` tuples
// at b * LS_CAP of the region arrays (in_hash / in_loc), in arbitrary order, and its output starts at start[b] (a scan of
// the cursors).  The index in an item is then only an arrival number, so equal hashes are put in order by their loc in
// the odd-even rounds (the loc gather of a tie: rare when hashes rarely repeat, which is when this path is taken).
#ifndef KS_BIN_TIEFIX
#define KS_BIN_TIEFIX 1  // target run: bin kernel 1.60 -> 1.46 ms (no hash-equality test in every compare of every round)
#endif
template <bool STEPPED, bool SCATTERED>
__global__ void __launch_bounds__(LS_THREADS, 3)
bucket_sort_bin_kernel(const uint64_t* __restrict__ in_hash, const uint64_t* __restrict__ in_loc,
                       uint64_t* __restrict__ out_hash, uint64_t* __restrict__ out_loc,
                       const uint32_t* __restrict__ start, const uint32_t* __restrict__ cursor, int lz, int tb,
                       uint64_t* __restrict__ counts, uint32_t* __restrict__ t_size, CsrOut f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* B = reinterpret_cast<uint64_t*>(smem_raw);            // [LS_CAP] items, grouped by bin
    uint32_t* cnt = reinterpret_cast<uint32_t*>(B + LS_CAP);        // [BN_BINS] counts, then exclusive offsets
    __shared__ uint32_t s_wsum[LS_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = blockIdx.x;  // buckets are independent: no ticket, no order
    reinterpret_cast<uint4*>(cnt)[tid] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(cnt)[tid + LS_THREADS] = make_uint4(0, 0, 0, 0);
    // a scattered bucket is a region of LS_CAP allocated entries at a position known without a load: the first half of
    // its rows is read before the size is known (what lies past the size is not used), so the size and the hashes come
    // back in one round trip instead of two (17 % of the kernel's stall samples sat on the dependent hash load: 1.46 -> 1.37 ms)
    constexpr int SPEC_ROWS = SCATTERED ? 6 : 0;  // 3072 of the 4096 slots: all the rows of an average bucket
    uint64_t item[8];
#pragma unroll
    for (int r = 0; r < SPEC_ROWS; r++) item[r] = in_hash[b * (uint32_t)LS_CAP + r * LS_THREADS + tid];
    const uint32_t n_oversize = f.oversize[0];  // (needed at the tail: on its way from here)
    const uint32_t s = start[b];
    const uint32_t s_in = SCATTERED ? b * (uint32_t)LS_CAP : s;
    const uint32_t m = SCATTERED ? min(cursor[b], (uint32_t)LS_CAP) : start[b + 1] - s;  // (an overflowing region: the host
                                                                                         // discards the build)
    if (m == 0) { if (tid == 0) counts[b] = 0; bucket_empty(b, s, f); return; }
    if (m > (uint32_t)LS_CAP) return;  // oversize: sorted and counted by the host-driven fallback
    const int sh = lz + tb;  // >= 12 on this path: the low 12 bits of (hash << sh) are zero and carry the index

    // 1. count: rows of 512 consecutive tuples, one per thread; the ticket a tuple draws is its slot in the bin
    // STEPPED (repeat-heavy input): a barrier after every row, so that a bin is ordered by row and only items of the
    // same row can be out of order -- the repeats of a hash then need a handful of moves instead of a full sort.
    uint32_t slot[4] = {0, 0, 0, 0};  // two 16-bit slots per word
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t j = r * LS_THREADS + tid;
        if (r * LS_THREADS >= m) break;  // uniform
        if (j < m) item[r] = (((r < SPEC_ROWS ? item[r] : in_hash[s_in + j]) << sh) & ~0xfffull) | j;
    }
    __syncthreads();  // the counters are zero
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t j = r * LS_THREADS + tid;
        if (r * LS_THREADS >= m) break;  // uniform
        if (j < m) slot[r >> 1] |= atomicAdd(&cnt[(uint32_t)(item[r] >> (64 - BN_BITS))], 1u) << (16 * (r & 1));
        if (STEPPED && (r + 1) * LS_THREADS < m) __syncthreads();  // uniform
    }
    __syncthreads();
    // 2. exclusive scan of the bin counts; thread t owns bins 8t .. 8t+7
    {
        uint4 c0 = reinterpret_cast<uint4*>(cnt)[2 * tid], c1 = reinterpret_cast<uint4*>(cnt)[2 * tid + 1];
        const uint32_t total = c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w;
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
        for (int w = 0; w < LS_WARPS; w++) off += w < (int)warp ? s_wsum[w] : 0u;
        uint4 e0, e1;
        e0.x = off + incl - total; e0.y = e0.x + c0.x; e0.z = e0.y + c0.y; e0.w = e0.z + c0.z;
        e1.x = e0.w + c0.w; e1.y = e1.x + c1.x; e1.z = e1.y + c1.y; e1.w = e1.z + c1.z;
        reinterpret_cast<uint4*>(cnt)[2 * tid] = e0;
        reinterpret_cast<uint4*>(cnt)[2 * tid + 1] = e1;
    }
    __syncthreads();
    // 3. scatter into the bins
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t j = r * LS_THREADS + tid;
        if (r * LS_THREADS >= m) break;  // uniform
        if (j < m) B[cnt[(uint32_t)(item[r] >> (64 - BN_BITS))] + ((slot[r >> 1] >> (16 * (r & 1))) & 0xffffu)] = item[r];
    }
    // the bucket's directory entries are prefixes of the bin index (dir_sub <= BN_BITS bits of the item's top): the first
    // tuple of entry x is the offset of bin x << (BN_BITS - dir_sub).  Picked up here, before the offsets are overwritten.
    DirFirst df;
    df.from_bins = n_oversize == 0 && f.dir_sub <= BN_BITS - 2;  // <= 2 entries per thread
    if (df.from_bins) {
        const int bs = BN_BITS - f.dir_sub;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint32_t x = tid + i * LS_THREADS;
            df.first[i] = x < (1u << f.dir_sub) ? cnt[x << bs] : m;
        }
    }
    __syncthreads();
    // 4. odd-even transposition until nothing moves.  The array is ordered by bin and, within a bin, by row; only items
    //    of the same (bin, row) can be out of place, so an item is a few positions from its final one and a handful of
    //    perfectly balanced rounds finish the sort -- for any input (a bucket full of one hash needs more rounds, not
    //    a different code path).
    {
        const uint32_t np0 = m >> 1, np1 = (m - 1) >> 1;  // pairs (2i, 2i+1) and (2i+1, 2i+2)
        ulonglong2* B2 = reinterpret_cast<ulonglong2*>(B);
        // x belongs after y: by item; with arrival numbers instead of ordered indices, equal hashes by their loc
        auto after = [&](uint64_t x, uint64_t y) -> bool {
#if KS_BIN_TIEFIX
            return x > y;  // (equal hashes are put in loc order after the rounds)
#else
            if (!SCATTERED || ((x ^ y) >> 12)) return x > y;
            return in_loc[s_in + (uint32_t)(x & 0xfffu)] > in_loc[s_in + (uint32_t)(y & 0xfffu)];
#endif
        };
        int again;
        do {
            int sw = 0;
            for (uint32_t i = tid; i < np0; i += LS_THREADS) {
                const ulonglong2 v = B2[i];
                if (after(v.x, v.y)) { B2[i] = make_ulonglong2(v.y, v.x); sw = 1; }
            }
            __syncthreads();
            for (uint32_t i = tid; i < np1; i += LS_THREADS) {
                const uint64_t x = B[2 * i + 1], y = B[2 * i + 2];
                if (after(x, y)) { B[2 * i + 1] = y; B[2 * i + 2] = x; sw = 1; }
            }
            again = __syncthreads_or(sw);
        } while (again);
#if KS_BIN_TIEFIX
        if (SCATTERED) {
            // equal hashes (rare on this path) sit next to each other in arrival order: the thread that owns the second
            // item of a run puts the run in loc order.  Runs are disjoint and a run's items differ in their low 12 bits
            // only, so the neighbour tests of other threads read the same hash bits whatever this thread has moved.
            for (uint32_t j = tid + 1; j < m; j += LS_THREADS) {
                const uint64_t it = B[j], pv = B[j - 1];
                if ((it ^ pv) >> 12) continue;
                if (j >= 2 && ((B[j - 2] ^ pv) >> 12) == 0) continue;  // not the second item of its run
                uint32_t e = j + 1;
                while (e < m && ((B[e] ^ it) >> 12) == 0) e++;
                for (uint32_t t = j; t < e; t++) {  // insertion by loc
                    const uint64_t x = B[t], lx = in_loc[s_in + (uint32_t)(x & 0xfffu)];
                    uint32_t u = t;
                    while (u > j - 1) {
                        const uint64_t y = B[u - 1];
                        if (in_loc[s_in + (uint32_t)(y & 0xfffu)] <= lx) break;
                        B[u] = y;
                        u--;
                    }
                    B[u] = x;
                }
            }
            __syncthreads();
        }
#endif
    }

    bucket_finish(B, cnt, m, s_in, s, b, lz, tb, in_loc, out_hash, out_loc, counts, t_size, f, n_oversize, df);
}

// scaled == 1: every complete window is a tuple; kept windows per protein straight from the offsets (the scattered
// input has no (protein, pos)-ordered tuple array to search)
__global__ void protein_windows_kernel(const uint64_t* __restrict__ offsets, uint32_t n_prot, uint32_t k,
                                       uint32_t* __restrict__ t_abund, uint32_t* __restrict__ t_size) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_prot) return;
    const uint64_t len = offsets[p + 1] - offsets[p];
    const uint32_t w = len >= k ? (uint32_t)(len - k + 1) : 0u;
    t_abund[p] = w;
    t_size[p] = w;
}

// Second level of the unstable partition: DS_TILE-sized chunks of the first-level regions into the final buckets.
// (64 registers, 4 CTAs per SM: 0.650 ms on the target run; forced to 48 registers / 5 CTAs it spills and takes 0.707 ms)
__global__ void __launch_bounds__(DS_THREADS)
pair_partition_kernel(const uint64_t* __restrict__ r1_hash, const uint64_t* __restrict__ r1_loc, uint32_t cap1,
                      const uint2* __restrict__ chunk_map, PairScatter sc) {
    __shared__ DenseScatterSmem s_sc;
    __shared__ uint64_t s_dk[DS_TILE], s_dv[DS_TILE];
    const uint2 e = chunk_map[blockIdx.x];  // (dense_chunks_kernel) region | pairs << 16, offset in the region
    const uint32_t nv = e.x >> 16, b1 = e.x & 0xffffu;
    if (nv == 0) return;
    const uint64_t base = (uint64_t)b1 * cap1 + e.y;
    uint64_t key[DS_ITEMS], val[DS_ITEMS];
    uint32_t valid = 0;
#pragma unroll
    for (int it = 0; it < DS_ITEMS; it++) {
        const uint32_t i = it * DS_THREADS + threadIdx.x;
        key[it] = val[it] = 0;
        if (i < nv) { key[it] = r1_hash[base + i]; val[it] = r1_loc[base + i]; valid |= 1u << it; }
    }
    scatter_pairs(key, valid, [&](int it) { return val[it]; }, sc, b1 << sc.bits, s_sc, s_dk, s_dv);
}

// Counts for ranges the bucket sort did not handle: oversize buckets (only_oversize = 1), or every range on the
// library-sort path (only_oversize = 0).
__global__ void __launch_bounds__(256)
range_count_kernel(const uint64_t* __restrict__ hash, const uint64_t* __restrict__ loc, const uint32_t* __restrict__ start,
                   int only_oversize, uint64_t* __restrict__ counts, uint32_t* __restrict__ t_size) {
    __shared__ uint32_t s_tk[8], s_tg[8];
    const uint32_t b = blockIdx.x;
    const uint32_t s = start[b], e = start[b + 1];
    if (only_oversize && e - s <= (uint32_t)LS_CAP) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t tk = 0, tg = 0;
    for (uint32_t i0 = s; i0 < e; i0 += 256) {
        const uint32_t i = i0 + tid;
        bool hk = false, hg = false;
        if (i < e) {
            const uint64_t h = hash[i];
            const uint32_t pid = (uint32_t)(loc[i] >> 32);
            if (i == 0) {
                hk = hg = true;
            } else {
                hk = h != hash[i - 1];
                hg = hk || pid != (uint32_t)(loc[i - 1] >> 32);
            }
            if (!hg) atomicSub(&t_size[pid], 1u);
        }
        tk += __popc(__ballot_sync(0xffffffffu, hk));
        tg += __popc(__ballot_sync(0xffffffffu, hg));
    }
    if (lane == 0) { s_tk[warp] = tk; s_tg[warp] = tg; }
    __syncthreads();
    if (tid == 0) {
        uint32_t k = 0, g = 0;
        for (int w = 0; w < 8; w++) { k += s_tk[w]; g += s_tg[w]; }
        counts[b] = (uint64_t)k | ((uint64_t)g << 32);
    }
}

// Exclusive scan of the per-range counts (both 32-bit halves at once), one CTA.  Also writes the totals and
// the two sentinel entries of the CSR arrays.
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const uint64_t* __restrict__ counts, uint32_t nb, uint64_t* __restrict__ prefix, uint64_t n,
                   uint64_t* __restrict__ d_counts, uint32_t* __restrict__ key_grp, uint32_t* __restrict__ grp_start) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + tid;
        const uint64_t v = i < nb ? counts[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_w[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if ((int)lane >= o) wi += t;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const uint64_t excl = s_carry + s_w[warp] + incl - v;
        if (i < nb) prefix[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        const uint64_t U = s_carry & 0xffffffffu, G = s_carry >> 32;
        prefix[nb] = s_carry;
        d_counts[0] = U;
        d_counts[1] = G;
        key_grp[U] = (uint32_t)G;
        grp_start[G] = (uint32_t)n;
    }
}

// One CTA per range: recompute the head flags of the sorted tuples and write the CSR arrays at the offsets
// the scan produced.  Ranges of any length are walked in chunks with a running base.
constexpr int CW_THREADS = 384;  // 384 x 8 = 3072 tuples per chunk: an average bucket in one chunk
constexpr int CW_ROWS = 8;
__global__ void __launch_bounds__(CW_THREADS)
csr_write_kernel(const uint64_t* __restrict__ hash, const uint64_t* __restrict__ loc, const uint32_t* __restrict__ start,
                 const uint64_t* __restrict__ prefix, uint64_t* __restrict__ keys, uint32_t* __restrict__ key_grp,
                 uint32_t* __restrict__ grp_start) {
    __shared__ uint32_t s_wk[CW_THREADS / 32], s_wg[CW_THREADS / 32];
    const uint32_t b = blockIdx.x;
    const uint32_t s = start[b], e = start[b + 1];
    if (s == e) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t pfx = prefix[b];
    uint32_t base_k = (uint32_t)pfx, base_g = (uint32_t)(pfx >> 32);
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t c0 = s; c0 < e; c0 += CW_THREADS * CW_ROWS) {
        uint64_t h[CW_ROWS];
        uint32_t bk[CW_ROWS], bg[CW_ROWS];
        uint32_t wk = 0, wg = 0;
#pragma unroll
        for (int r = 0; r < CW_ROWS; r++) {
            const uint32_t i = c0 + (warp * CW_ROWS + r) * 32 + lane;
            bool hk = false, hg = false;
            h[r] = 0;
            uint32_t pid = 0;
            if (i < e) { h[r] = hash[i]; pid = (uint32_t)(loc[i] >> 32); }
            uint64_t hp = __shfl_up_sync(0xffffffffu, h[r], 1);
            uint32_t pp = __shfl_up_sync(0xffffffffu, pid, 1);
            if (i < e) {
                if (i == 0) {
                    hk = hg = true;
                } else {
                    if (lane == 0) { hp = hash[i - 1]; pp = (uint32_t)(loc[i - 1] >> 32); }
                    hk = h[r] != hp;
                    hg = hk || pid != pp;
                }
            }
            bk[r] = __ballot_sync(0xffffffffu, hk);
            bg[r] = __ballot_sync(0xffffffffu, hg);
            wk += __popc(bk[r]);
            wg += __popc(bg[r]);
        }
        if (lane == 0) { s_wk[warp] = wk; s_wg[warp] = wg; }
        __syncthreads();
        uint32_t rk = base_k, rg = base_g;
#pragma unroll
        for (int w = 0; w < CW_THREADS / 32; w++) {
            if (w < (int)warp) { rk += s_wk[w]; rg += s_wg[w]; }
            base_k += s_wk[w];
            base_g += s_wg[w];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < CW_ROWS; r++) {
            const uint32_t i = c0 + (warp * CW_ROWS + r) * 32 + lane;
            const uint32_t g = rg + __popc(bg[r] & lt);
            if ((bg[r] >> lane) & 1u) grp_start[g] = i;
            if ((bk[r] >> lane) & 1u) {
                const uint32_t u = rk + __popc(bk[r] & lt);
                keys[u] = h[r];
                key_grp[u] = g;
            }
            rk += __popc(bk[r]);
            rg += __popc(bg[r]);
        }
    }
}

// dir[b] = index of the first key whose bucket is >= b; dir[2^bits] = U.  Grid-stride: U is only known on the device.
__global__ void dir_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ d_counts,
                           uint32_t* __restrict__ dir, int bits, int shift) {
    const uint64_t U = d_counts[0];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint32_t nb = 1u << bits;
    if (U == 0) {
        for (uint64_t b = i0; b <= nb; b += stride) dir[b] = 0;
        return;
    }
    for (uint64_t i = i0; i < U; i += stride) {
        const uint32_t b = (uint32_t)(keys[i] >> shift);
        const int64_t bp = i ? (int64_t)(uint32_t)(keys[i - 1] >> shift) : -1;
        for (int64_t x = bp + 1; x <= (int64_t)b; x++) dir[x] = (uint32_t)i;
    }
    // buckets above the last key (a large share when max_hash is not a power of two): filled by everybody
    const uint32_t b_last = (uint32_t)(keys[U - 1] >> shift);
    for (uint64_t x = (uint64_t)b_last + 1 + i0; x <= nb; x += stride) dir[x] = (uint32_t)U;
}

int msd_top_bits(uint64_t n, int lz, uint64_t max_hash, double avg_postings) {
    // a k-mer space so small that its frequent hashes fill a bucket on their own (dayhoff k <= 7 at 10 M residues: 36 postings
    // per hash on average, far more for the frequent ones): buckets would be oversize and take the per-range fallback;
    // the library sort of all bits is the fast path there (measured: dayhoff k=6 56 ms -> 1.3 ms)
    if (avg_postings > 16.0) return -1;
    // smallest tb whose average bucket is <= 3072 tuples (LS_CAP = 4096 leaves > 18 sigma for uniform hashes).
    // Hashes fill only [0, max_hash] of the 2^(64-lz) range the buckets span, so the used buckets are fuller.
    const double fill = lz >= 64 ? 1.0 : ((double)max_hash + 1.0) / std::ldexp(1.0, 64 - lz);
    int tb = 0;
    while (tb < 24 && (double)n / (std::ldexp(1.0, tb) * fill) > 3072.0) tb++;
    if ((double)n / (std::ldexp(1.0, tb) * fill) > 3072.0) return -1;
    // the item layout needs lz + tb >= 12 (the index in the bucket takes the low 12 bits of the shifted hash): small
    // inputs simply get more, smaller buckets -- still two partition passes instead of eight
    if (lz + tb < 12) tb = 12 - lz;
    if (n < (1u << 16)) return -1;  // tiny: the library sort of all bits is as fast as anything
    return tb;
}

cudaError_t library_sort(uint64_t* hash_a, uint64_t* loc_a, uint64_t* hash_b, uint64_t* loc_b, uint64_t n, int begin_bit,
                         int end_bit, void* temp, size_t temp_bytes, cudaStream_t stream, int* out_in_a,
                         uint64_t* n_launches) {
    cub::DoubleBuffer<uint64_t> k(hash_a, hash_b), v(loc_a, loc_b);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, (int64_t)n, begin_bit, end_bit, stream);
    *out_in_a = k.Current() == hash_a ? 1 : 0;
    if (n_launches) *n_launches += 2 + (end_bit - begin_bit + 7) / 8;  // histogram + scan + one onesweep pass per 8 bits
    return e;
}

uint64_t max_ranges(uint64_t n) {
    uint64_t nb = std::max<uint64_t>(n / 512 + 2, 4096);  // msd_top_bits: average used bucket > 1536, fill > 0.5
    nb = std::min<uint64_t>(nb, 1ull << 24);
    return std::max<uint64_t>(nb, n / LS_CAP + 2);
}

size_t table_bytes(uint64_t n) {
    // start[nb+1] + oversize[2] (u32), counts[nb] + prefix[nb+1] + status[nb] (u64)
    const uint64_t nb = max_ranges(n);
    return (size_t)(((nb + 8) * 4 + (3 * nb + 6) * 8 + 1023) & ~(size_t)255);
}

// hash[i] of the sorted tuples from the CSR arrays: one thread per key writes its row.
__global__ void expand_hash_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ key_grp,
                                   const uint32_t* __restrict__ grp_start, const uint64_t* __restrict__ d_counts,
                                   uint64_t* __restrict__ hash) {
    const uint64_t U = d_counts[0];
    for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < U; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = keys[u];
        const uint32_t e = grp_start[key_grp[u + 1]];
        for (uint32_t i = grp_start[key_grp[u]]; i < e; i++) hash[i] = h;
    }
}

// the same over the segmented layout: one CTA per sort bucket
__global__ void expand_hash_seg_kernel(CsrView v, uint64_t* __restrict__ hash) {
    const uint32_t b = blockIdx.x;
    const uint32_t kb = v.seg_start[b] + b, tk = (uint32_t)v.seg_counts[b];
    for (uint32_t i = threadIdx.x; i < tk; i += blockDim.x) {
        const uint32_t u = kb + i;
        const uint64_t h = v.keys[u];
        const uint32_t e = v.grp_start[v.key_grp[u + 1]];
        for (uint32_t t = v.grp_start[v.key_grp[u]]; t < e; t++) hash[t] = h;
    }
}

}  // namespace

cudaError_t launch_directory(const uint64_t* keys, const uint64_t* d_counts, uint32_t* dir, int dir_bits, int dir_shift,
                             cudaStream_t stream) {
    dir_kernel<<<148 * 8, 256, 0, stream>>>(keys, d_counts, dir, dir_bits, dir_shift);
    return cudaGetLastError();
}

cudaError_t expand_sorted_hash(const CsrView& v, uint64_t* hash, cudaStream_t stream) {
    if (v.n == 0) return cudaSuccess;
    if (v.dir_sub != DIR_SUB_COMPACT) expand_hash_seg_kernel<<<v.seg_nb, 256, 0, stream>>>(v, hash);
    else expand_hash_kernel<<<148 * 8, 256, 0, stream>>>(v.keys, v.key_grp, v.grp_start, v.d_counts, hash);
    return cudaGetLastError();
}

int build_top_bits(uint64_t n, int end_bit, uint64_t max_hash, double avg_postings) {
    return n ? msd_top_bits(n, 64 - end_bit, max_hash, avg_postings) : -1;
}
uint64_t build_slack(uint64_t n) { return max_ranges(n) + 2; }

PairSortPlan pair_sort_plan(uint64_t n, int end_bit, uint64_t max_hash) {
    PairSortPlan p;
    const int lz = 64 - end_bit;
    const int tb = msd_top_bits(n, lz, max_hash, 0.0);
    if (tb < 0 || tb > 2 * DS_MAX_BITS) return p;
    p.custom = 1;
    p.total = tb;
    p.l1 = tb < DS_MAX_BITS ? tb : DS_MAX_BITS;
    p.l2 = tb - p.l1;
    const double fill = lz >= 64 ? 1.0 : ((double)max_hash + 1.0) / std::ldexp(1.0, 64 - lz);  // used share of the bins
    const double per1 = (double)n / (std::ldexp(1.0, p.l1) * fill);
    p.cap1 = p.l2 ? (uint32_t)(per1 + per1 / 8 + 16384) : (uint32_t)LS_CAP;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    p.off_r1_hash = take(((size_t)p.cap1 << p.l1) * 8);
    p.off_r1_loc = take(((size_t)p.cap1 << p.l1) * 8);
    p.off_r2_hash = p.l2 ? take(((size_t)LS_CAP << tb) * 8) : p.off_r1_hash;
    p.off_r2_loc = p.l2 ? take(((size_t)LS_CAP << tb) * 8) : p.off_r1_loc;
    p.off_small = off;
    p.off_cursor1 = take(((size_t)1 << p.l1) * 4);
    p.off_cursor2 = p.l2 ? take(((size_t)1 << tb) * 4) : p.off_cursor1;
    p.off_overflow = take(8);
    p.small_bytes = off - p.off_small;  // zeroed before the sketch kernel scatters into the regions
    // chunks of the second level: n is exact for scaled == 1 (every region's last chunk may be partial); for scaled > 1 it is
    // an estimate, and the bound is what the regions can hold
    p.max_chunks = max_hash == ~0ull ? (uint32_t)(n / DS_TILE + ((size_t)1 << p.l1) + 1)
                                     : (uint32_t)((((size_t)p.cap1 + DS_TILE - 1) / DS_TILE) << p.l1);
    p.off_chunks = take((size_t)p.max_chunks * 8);
    p.off_bstart = take((((size_t)1 << tb) + 1) * 4);
    p.bytes = off;
    return p;
}

size_t build_temp_bytes(uint64_t n, int end_bit) {
    size_t a = 0, b = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, v, (int64_t)n, 0, end_bit);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint64_t*)nullptr,
                                    (uint64_t*)nullptr, (int64_t)n, 0, end_bit);
    return table_bytes(n) + (a > b ? a : b) + 512;
}

cudaError_t build_index(const BuildArgs& a, cudaStream_t stream, int* out_in_a, uint64_t* sort_launches,
                        uint64_t* csr_launches) {
#define KS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
    const uint64_t n = a.n;
    const int lz = 64 - a.end_bit;
    *out_in_a = 1;
    if (a.hash_written) *a.hash_written = 1;
    if (a.overflow_dev) *a.overflow_dev = nullptr;
    if (a.out_dir_sub) *a.out_dir_sub = DIR_SUB_COMPACT;
    if (a.out_seg_nb) *a.out_seg_nb = 0;
    if (a.out_seg_start) *a.out_seg_start = nullptr;
    if (a.out_seg_counts) *a.out_seg_counts = nullptr;
    if (a.plan.custom && n) {
        // scattered input: [second scatter level,] bucket offsets, bin kernel straight from the final buckets
        const PairSortPlan& pl = a.plan;
        char* w = (char*)a.work;
        const uint32_t nb = 1u << pl.total;
        uint32_t* cursor1 = (uint32_t*)(w + pl.off_cursor1);
        uint32_t* cursor2 = (uint32_t*)(w + pl.off_cursor2);
        uint32_t* overflow = (uint32_t*)(w + pl.off_overflow);
        uint2* chunk_map = (uint2*)(w + pl.off_chunks);
        uint32_t* bstart = (uint32_t*)(w + pl.off_bstart);
        const uint64_t* r2h = (const uint64_t*)(w + pl.off_r2_hash);
        const uint64_t* r2l = (const uint64_t*)(w + pl.off_r2_loc);
        if (a.n_prot) {
            if (a.abund_ready)  // distinct hashes per protein start from the kept windows the sketch kernel counted
                KS_TRY(cudaMemcpyAsync(a.t_size, a.t_abund, (size_t)a.n_prot * 4, cudaMemcpyDeviceToDevice, stream));
            else
                protein_windows_kernel<<<(a.n_prot + 255) / 256, 256, 0, stream>>>(a.offsets, a.n_prot, a.k, a.t_abund, a.t_size);
        }
        if (pl.l2) {
            dense_chunks_kernel<<<1, 256, 0, stream>>>(cursor1, 1u << pl.l1, pl.cap1, chunk_map, pl.max_chunks);
            PairScatter sc;
            sc.out_key = (uint64_t*)(w + pl.off_r2_hash); sc.out_val = (uint64_t*)(w + pl.off_r2_loc); sc.cursor = cursor2;
            sc.cap = (uint32_t)LS_CAP; sc.shift = 64 - pl.total; sc.bits = pl.l2; sc.lz = lz; sc.overflow = overflow;
            pair_partition_kernel<<<pl.max_chunks, DS_THREADS, 0, stream>>>((const uint64_t*)(w + pl.off_r1_hash),
                                                                           (const uint64_t*)(w + pl.off_r1_loc), pl.cap1, chunk_map, sc);
            KS_TRY(cudaGetLastError());
        }
        dense_bucket_offsets_kernel<<<1, 1024, 0, stream>>>(cursor2, nb, (uint32_t)LS_CAP, bstart);
        if (a.ev_partitioned) KS_TRY(cudaEventRecord(a.ev_partitioned, stream));
        // tables of the bucket sort (oversize word stays 0: the CSR write is always fused here)
        char* tp = (char*)a.temp;
        uint32_t* oversize = (uint32_t*)tp + nb + 1;
        uint64_t* counts = (uint64_t*)(tp + (((size_t)(nb + 8) * 4 + 7) & ~(size_t)7));
        KS_TRY(cudaMemsetAsync(oversize, 0, 8, stream));
        KS_TRY(cudaMemsetAsync(a.d_counts, 0, 16, stream));
        if (a.dir_bits < pl.total) return cudaErrorInvalidValue;  // (the caller sizes the directory from build_top_bits)
        CsrOut f;
        f.oversize = oversize; f.keys = a.keys; f.key_grp = a.key_grp; f.grp_start = a.grp_start;
        f.d_counts = (unsigned long long*)a.d_counts; f.nb = nb;
        f.dir = a.dir; f.dir_sub = a.dir_bits - pl.total;
        KS_TRY(cudaFuncSetAttribute(bucket_sort_bin_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BN_SMEM));
        bucket_sort_bin_kernel<false, true><<<nb, LS_THREADS, BN_SMEM, stream>>>(r2h, r2l, a.hash_b, a.loc_a, bstart, cursor2, lz, pl.total,
                                                                                 counts, a.t_size, f);
        KS_TRY(cudaGetLastError());
        *sort_launches += pl.l2 ? 5 : 3;
        if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
        if (a.overflow_dev) *a.overflow_dev = overflow;  // read by the caller together with the totals: no round trip here
        if (a.hash_written) *a.hash_written = 0;
        if (a.out_dir_sub) *a.out_dir_sub = f.dir_sub;
        if (a.out_seg_nb) *a.out_seg_nb = nb;
        if (a.out_seg_start) *a.out_seg_start = bstart;
        if (a.out_seg_counts) *a.out_seg_counts = counts;
        *out_in_a = 1;
        return cudaGetLastError();
    }
    if (a.n_prot) {
        protein_abund_kernel<<<(a.n_prot + 255) / 256, 256, 0, stream>>>(a.loc_a, n, a.n_prot, a.t_abund, a.t_size);
        KS_TRY(cudaGetLastError());
        *csr_launches += 1;
    }
    if (n == 0) {
        KS_TRY(cudaMemsetAsync(a.d_counts, 0, 16, stream));
        KS_TRY(cudaMemsetAsync(a.key_grp, 0, 4, stream));
        KS_TRY(cudaMemsetAsync(a.grp_start, 0, 4, stream));
        if (a.ev_partitioned) KS_TRY(cudaEventRecord(a.ev_partitioned, stream));
        if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
        dir_kernel<<<1, 256, 0, stream>>>(a.keys, a.d_counts, a.dir, a.dir_bits, a.dir_shift);
        *csr_launches += 1;
        return cudaGetLastError();
    }
    // carve temp: [tables | library scratch]
    const size_t tbytes = table_bytes(n);
    char* tp = (char*)a.temp;
    void* lib_temp = tp + tbytes;
    size_t lib_bytes = a.temp_bytes - tbytes;
    const int tb = msd_top_bits(n, lz, a.max_hash, a.avg_postings);
    const uint32_t nb = tb >= 0 ? (1u << tb) : (uint32_t)((n + LS_CAP - 1) / LS_CAP);
    uint32_t* start = (uint32_t*)tp;                                             // [nb + 1]
    uint32_t* oversize = start + nb + 1;                                          // [2]
    uint64_t* counts = (uint64_t*)(tp + (((size_t)(nb + 8) * 4 + 7) & ~(size_t)7));  // [nb]
    uint64_t* prefix = counts + nb;                                               // [nb + 1]

    const uint64_t *fh, *fl;  // final sorted tuples
    if (tb < 0) {
        int in_a = 1;
        KS_TRY(library_sort(a.hash_a, a.loc_a, a.hash_b, a.loc_b, n, 0, a.end_bit, lib_temp, lib_bytes, stream, &in_a, sort_launches));
        *out_in_a = in_a;
        fh = in_a ? a.hash_a : a.hash_b;
        fl = in_a ? a.loc_a : a.loc_b;
        if (a.ev_partitioned) KS_TRY(cudaEventRecord(a.ev_partitioned, stream));
        if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
        fixed_ranges_kernel<<<(nb + 1 + 255) / 256, 256, 0, stream>>>(n, nb, start);
        range_count_kernel<<<nb, 256, 0, stream>>>(fh, fl, start, 0, counts, a.t_size);
        KS_TRY(cudaGetLastError());
        *csr_launches += 2;
    } else {
        // 1. partition by the top tb bits (stable): library onesweep passes over those bits only
        int in_a = 1;
        if (tb > 0) KS_TRY(library_sort(a.hash_a, a.loc_a, a.hash_b, a.loc_b, n, a.end_bit - tb, a.end_bit, lib_temp, lib_bytes, stream, &in_a, sort_launches));
        if (a.ev_partitioned) KS_TRY(cudaEventRecord(a.ev_partitioned, stream));
        uint64_t* sh = in_a ? a.hash_a : a.hash_b;
        uint64_t* sl = in_a ? a.loc_a : a.loc_b;
        uint64_t* dh = in_a ? a.hash_b : a.hash_a;
        uint64_t* dl = in_a ? a.loc_b : a.loc_a;
        // 2. bucket boundaries; one CTA per bucket sorts it in shared memory and counts its heads
        bucket_start_kernel<<<(nb + 1 + 255) / 256, 256, 0, stream>>>(sh, n, lz, tb, start, oversize);
        oversize_count_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(start, nb, oversize);
        if (a.dir_bits < tb) return cudaErrorInvalidValue;  // (the caller sizes the directory from build_top_bits)
        CsrOut f;
        f.oversize = oversize; f.keys = a.keys; f.key_grp = a.key_grp; f.grp_start = a.grp_start;
        f.d_counts = (unsigned long long*)a.d_counts; f.nb = nb;
        f.dir = a.dir; f.dir_sub = a.dir_bits - tb;
        KS_TRY(cudaMemsetAsync(a.d_counts, 0, 16, stream));
        // repeat-heavy inputs (small k-mer space, e.g. hp k=24) take the two-pass stable variant, everything else the bin
        // variant; ls_variant (KS_LS_VARIANT = rep | bn | bs: bin variant without / with a barrier per row) is a test hook
        const bool bin = a.ls_variant ? a.ls_variant >= 2 : (a.repeat_heavy == 0);
        if (bin) {
            KS_TRY(cudaFuncSetAttribute(bucket_sort_bin_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BN_SMEM));
            KS_TRY(cudaFuncSetAttribute(bucket_sort_bin_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BN_SMEM));
            const bool stepped = a.ls_variant == 3;
            if (stepped) bucket_sort_bin_kernel<true, false><<<nb, LS_THREADS, BN_SMEM, stream>>>(sh, sl, dh, dl, start, nullptr, lz, tb, counts, a.t_size, f);
            else bucket_sort_bin_kernel<false, false><<<nb, LS_THREADS, BN_SMEM, stream>>>(sh, sl, dh, dl, start, nullptr, lz, tb, counts, a.t_size, f);
        } else {
            KS_TRY(cudaFuncSetAttribute(bucket_sort_rep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM_REP));
            bucket_sort_rep_kernel<<<nb, LS_THREADS, LS_SMEM_REP, stream>>>(sh, sl, dh, dl, start, lz, tb, counts, a.t_size, f);
        }
        KS_TRY(cudaGetLastError());
        *sort_launches += 3;
        // 3. oversize buckets (heavy repeats of few hashes): library sort of the remaining bits, range by range
        uint32_t n_over = 0;
        KS_TRY(cudaMemcpyAsync(&n_over, oversize, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        KS_TRY(cudaStreamSynchronize(stream));
        if (n_over) {
            std::vector<uint32_t> st(nb + 1);
            KS_TRY(cudaMemcpyAsync(st.data(), start, (size_t)(nb + 1) * 4, cudaMemcpyDeviceToHost, stream));
            KS_TRY(cudaStreamSynchronize(stream));
            for (uint32_t b = 0; b < nb; b++) {
                const uint64_t s0 = st[b], m = st[b + 1] - st[b];
                if (m <= (uint64_t)LS_CAP) continue;
                KS_TRY(cub::DeviceRadixSort::SortPairs(lib_temp, lib_bytes, sh + s0, dh + s0, sl + s0, dl + s0, (int64_t)m, 0,
                                                       a.end_bit - tb, stream));
                *sort_launches += 2 + (a.end_bit - tb + 7) / 8;
            }
            range_count_kernel<<<nb, 256, 0, stream>>>(dh, dl, start, 1, counts, a.t_size);
            KS_TRY(cudaGetLastError());
            *sort_launches += 1;
        }
        *out_in_a = in_a ? 0 : 1;
        fh = dh;
        fl = dl;
        if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
        if (n_over == 0) {  // the bucket sort wrote keys / key_grp / grp_start and the directory itself (segmented layout)
            if (a.hash_written) *a.hash_written = 0;  // ... and skipped the sorted hash column
            if (a.out_dir_sub) *a.out_dir_sub = f.dir_sub;
            if (a.out_seg_nb) *a.out_seg_nb = nb;
            if (a.out_seg_start) *a.out_seg_start = start;
            if (a.out_seg_counts) *a.out_seg_counts = counts;
            return cudaGetLastError();
        }
    }
    // 4. scan of the range counts, CSR write, directory
    scan_counts_kernel<<<1, 1024, 0, stream>>>(counts, nb, prefix, n, a.d_counts, a.key_grp, a.grp_start);
    csr_write_kernel<<<nb, CW_THREADS, 0, stream>>>(fh, fl, start, prefix, a.keys, a.key_grp, a.grp_start);
    dir_kernel<<<148 * 8, 256, 0, stream>>>(a.keys, a.d_counts, a.dir, a.dir_bits, a.dir_shift);
    *csr_launches += 3;
    return cudaGetLastError();
#undef KS_TRY
}

}  // namespace ks
