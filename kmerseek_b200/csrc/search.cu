// Query kernels for sm_100a: batched search of query sketches in the CSR index -- per-(query, target) aggregation, fp64
// scores, the hit list (query pos x target pos) -- and the multi-GPU merge of shard results.
//
// Replaces sourmash_plugin_branchwater.do_manysearch as called at src/python/kmerseek/search.py:125-141
// (all-pairs sorted-list intersections) and the polars join on (encoded, hashval) at search.py:204-213.
// Score formulae: SURVEY.md Appendix A.6.
//
// Shape of the hand-written path (queries of up to QK_MAX_WINDOWS windows): ONE CTA PER QUERY does everything a query
// needs in shared memory --
//   translate + MurmurHash3 of every window, FracMinHash filter                 (the sketch of the query)
//   bitonic sort of the kept hashes, run heads -> sorted distinct mins + abundances
//   one directory + binary-search lookup per distinct hash
//   (target, abundance) records of the found keys' groups, expanded by a load-balanced search over the scanned group
//     counts, in windows of the protein-id space that fit shared memory (normally one window = everything)
//   bitonic sort of the records by (target, abundance); run heads -> one pair per target: |Q n T|, sum, median and
//     population deviation of the abundances (median needs the abundance order)
//   pairs are staged compactly through one atomic reservation per window
// -- then one single-CTA scan over the queries gives every query its pair / hit / sketch offsets and the totals the host
// reads (the only round trip before the result block is sized), a streaming kernel turns staged pairs into the 19 result
// columns at their (query, target) position, and the hit list is expanded by a search over (query, window) offsets.
// No library sort or scan on this path.  Queries longer than that take the round-1 pipeline (search_device_legacy).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.cuh"
#include "search.cuh"

namespace ks {

namespace {

constexpr int TB = 256;
inline unsigned blocks_for(uint64_t n) { return (unsigned)((n + TB - 1) / TB); }
inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

inline int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b)) b++;
    return b;
}

// ---- lookups ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_key(const CsrView& c, uint64_t h) {
    const uint64_t x = h >> c.dir_shift;
    if (x >= (1ull << c.dir_bits)) return 0xffffffffu;
    const uint64_t b = x + (x >> c.dir_sub);  // segmented layout: every sort bucket has one more entry (its end); compact: + 0
    uint32_t lo = c.dir[b], hi = c.dir[b + 1];
    const uint32_t end = hi;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (c.keys[mid] < h) lo = mid + 1; else hi = mid;
    }
    return (lo < end && c.keys[lo] == h) ? lo : 0xffffffffu;
}

__device__ __forceinline__ uint32_t group_pid(const CsrView& c, uint32_t g) { return (uint32_t)(c.loc[c.grp_start[g]] >> 32); }

// ---- block-level primitives of the query kernel ------------------------------------------------
// In-place exclusive scan of a[0 .. n) (shared memory); a[n] <- total; returns the total.  All threads call; a[] must be
// complete (barrier before).  Ends with a barrier.
template <int T>
__device__ __forceinline__ uint32_t block_scan(uint32_t* a, uint32_t n, uint32_t* s_warp) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n + T - 1) / T;
    const uint32_t b = min(tid * per, n), e = min(b + per, n);
    uint32_t sum = 0;
    for (uint32_t i = b; i < e; i++) sum += a[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < T / 32; w++) {
        const uint32_t t = s_warp[w];
        if (w < (int)warp) woff += t;
        total += t;
    }
    uint32_t run = woff + incl - sum;
    for (uint32_t i = b; i < e; i++) { const uint32_t v = a[i]; a[i] = run; run += v; }
    if (tid == 0) a[n] = total;
    __syncthreads();
    return total;
}

// Sum over the block of a 64-bit value (every thread gets it).  Ends with a barrier.
template <int T>
__device__ __forceinline__ uint64_t block_sum64(uint64_t v, uint64_t* s_w64) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // s_w64 may still be read from the previous call
    if (lane == 0) s_w64[warp] = v;
    __syncthreads();
    uint64_t t = 0;
#pragma unroll
    for (int w = 0; w < T / 32; w++) t += s_w64[w];
    return t;
}

template <int T>
__device__ __forceinline__ uint32_t block_min32(uint32_t v, uint32_t* s_warp) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    uint32_t t = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < T / 32; w++) t = min(t, s_warp[w]);
    return t;
}

// Bitonic sort of key[0 .. n_pad) ascending (n_pad a power of two >= 2; the caller pads with ~0), with an optional 16-bit
// payload.  key[] must be complete (barrier before).  Ends with a barrier.
template <int T, bool PAYLOAD>
__device__ __forceinline__ void bitonic_sort(uint64_t* key, uint16_t* val, uint32_t n_pad) {
    for (uint32_t k = 2; k <= n_pad; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < (n_pad >> 1); i += T) {
                const uint32_t l = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const uint32_t r = l | j;
                const uint64_t a = key[l], b = key[r];
                const bool asc = (l & k) == 0;
                if (asc ? a > b : a < b) {
                    key[l] = b; key[r] = a;
                    if (PAYLOAD) { const uint16_t t = val[l]; val[l] = val[r]; val[r] = t; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint32_t pow2_at_least(uint32_t n) {
    uint32_t p = 2;
    while (p < n) p <<= 1;
    return p;
}

struct QueryKernelArgs {
    const uint8_t* res;
    const uint64_t* offs;
    uint32_t nq, k;
    uint64_t max_hash;
    CsrView c;
    QueryScratch s;
    int want_hits;
    uint32_t w_lo, w_hi;  // this launch takes the queries with w_lo < windows <= w_hi (and the empty ones when w_lo == 0)
};

template <int QCAP, int RCAP>
constexpr size_t query_smem_bytes() {
    constexpr int KCAP = QCAP > RCAP ? QCAP : RCAP;
    return (size_t)KCAP * 8 + (size_t)(KCAP + 1) * 4 + 3 * (size_t)(QCAP + 1) * 4 + 2 * (size_t)QCAP * 2 + (size_t)QCAP + SK_MAX_K + 16;
}

template <int QCAP, int RCAP, int T>
__global__ void __launch_bounds__(T) query_kernel(QueryKernelArgs a, Lut256 lut) {
    constexpr int KCAP = QCAP > RCAP ? QCAP : RCAP;
    static_assert(QCAP <= RCAP, "the records of one target (at most one per distinct query hash) must fit a window");
    extern __shared__ __align__(16) unsigned char q_smem[];
    uint64_t* s_key = reinterpret_cast<uint64_t*>(q_smem);    // [KCAP] hashes, then records
    uint32_t* A = reinterpret_cast<uint32_t*>(s_key + KCAP);  // [KCAP + 1] scan space
    uint32_t* B = A + KCAP + 1;                               // [QCAP + 1] first sorted slot of an entry; later: end group
    uint32_t* Cg = B + QCAP + 1;                              // [QCAP + 1] per entry: key index; later: current group
    uint32_t* D = Cg + QCAP + 1;                              // [QCAP + 1] per entry: end group of the current window
    uint16_t* s_idx = reinterpret_cast<uint16_t*>(D + QCAP + 1);  // [QCAP] window of a sorted slot
    uint16_t* s_went = s_idx + QCAP;                          // [QCAP] entry of a window (0xffff: not kept)
    uint8_t* s_res = reinterpret_cast<uint8_t*>(s_went + QCAP);  // [QCAP + SK_MAX_K] translated residues
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_warp[T / 32];
    __shared__ uint64_t s_w64[T / 32];
    __shared__ uint32_t s_n;
    __shared__ unsigned long long s_base;

    const uint32_t tid = threadIdx.x;
    const uint32_t q = blockIdx.x;
    const uint64_t o = a.offs[q];
    const uint64_t len64 = a.offs[q + 1] - o;
    const uint32_t k = a.k;
    const uint64_t W64 = len64 >= k ? len64 - k + 1 : 0;
    if (W64 > a.w_hi || (W64 <= a.w_lo && !(a.w_lo == 0 && W64 == 0))) return;  // another launch's query (uniform)
    const uint32_t W = (uint32_t)W64, len = (uint32_t)len64;
    if (W == 0) {
        if (tid == 0) { a.s.e_count[q] = 0; a.s.p_count[q] = 0; a.s.h_count[q] = 0; }
        return;
    }
    // ---- 1. the query's sketch: translate, hash, filter ----------------------------------------------------------
    for (uint32_t i = tid; i < 256; i += T) s_lut[i] = lut.b[i];
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (uint32_t i = tid; i < len; i += T) s_res[i] = s_lut[a.res[o + i]];
    for (uint32_t w = tid; w < W; w += T) s_went[w] = 0xffffu;
    __syncthreads();
    for (uint32_t w = tid; w < W; w += T) {
        const uint64_t h = murmur_bytes(s_res + w, k);
        if (h != 0 && h <= a.max_hash) {
            const uint32_t slot = atomicAdd(&s_n, 1u);
            s_key[slot] = h;
            s_idx[slot] = (uint16_t)w;
        }
    }
    __syncthreads();
    const uint32_t n = s_n;
    if (n == 0) {  // nothing kept (scaled > 1 on a short query)
        if (tid == 0) { a.s.e_count[q] = 0; a.s.p_count[q] = 0; a.s.h_count[q] = 0; }
        if (a.want_hits) for (uint32_t w = tid; w < W; w += T) { a.s.win_key[o + w] = 0xffffffffu; a.s.win_hoff[o + w] = 0; }
        return;
    }
    const uint32_t n_pad_q = pow2_at_least(n);
    for (uint32_t i = n + tid; i < n_pad_q; i += T) { s_key[i] = ~0ull; s_idx[i] = 0xffffu; }
    __syncthreads();
    bitonic_sort<T, true>(s_key, s_idx, n_pad_q);
    // ---- 2. distinct hashes (entries), abundances --------------------------------------------------------------------
    for (uint32_t i = tid; i < n; i += T) A[i] = (i == 0 || s_key[i] != s_key[i - 1]) ? 1u : 0u;
    __syncthreads();
    const uint32_t E = block_scan<T>(A, n, s_warp);
    for (uint32_t i = tid; i < n_pad_q; i += T) {
        if (i < n) {
            const bool head = i == 0 || s_key[i] != s_key[i - 1];
            const uint32_t e = A[i] + (head ? 1u : 0u) - 1u;
            if (s_idx[i] != 0xffffu) s_went[s_idx[i]] = (uint16_t)e;
            if (head) B[e] = i;
        } else if (s_idx[i] != 0xffffu) {
            // a real hash of exactly 2^64 - 1 that the sort left behind a pad (equal keys): it belongs to the last entry
            s_went[s_idx[i]] = (uint16_t)(E - 1);
        }
    }
    if (tid == 0) B[E] = n;
    __syncthreads();
    // ---- 3. one lookup per entry -------------------------------------------------------------------------------------
    for (uint32_t e = tid; e < E; e += T) {
        const uint64_t h = s_key[B[e]];
        a.s.ent_hash[o + e] = h;
        a.s.ent_abund[o + e] = B[e + 1] - B[e];
        Cg[e] = find_key(a.c, h);
    }
    if (tid == 0) a.s.e_count[q] = E;
    __syncthreads();
    // ---- 4. hits: per window the key of its hash and the number of hits before it (query order) ---------------------
    if (a.want_hits) {
        for (uint32_t e = tid; e < E; e += T) {
            const uint32_t u = Cg[e];
            D[e] = u == 0xffffffffu ? 0u : a.c.grp_start[a.c.key_grp[u + 1]] - a.c.grp_start[a.c.key_grp[u]];
        }
        __syncthreads();
        uint64_t mine = 0;
        for (uint32_t w = tid; w < W; w += T) {
            const uint32_t e = s_went[w];
            const uint32_t l = e == 0xffffu ? 0u : D[e];
            A[w] = l;
            mine += l;
        }
        const uint64_t H = block_sum64<T>(mine, s_w64);  // (its barriers also complete A[])
        block_scan<T>(A, W, s_warp);
        for (uint32_t w = tid; w < W; w += T) {
            const uint32_t e = s_went[w];
            a.s.win_key[o + w] = e == 0xffffu ? 0xffffffffu : Cg[e];
            a.s.win_hoff[o + w] = A[w];
        }
        if (tid == 0) {
            a.s.h_count[q] = H;
            if (H >> 32) atomicOr(reinterpret_cast<unsigned long long*>(a.s.totals + QT_FLAGS), (unsigned long long)QF_HITS_OVERFLOW);
        }
        __syncthreads();
    } else if (tid == 0) {
        a.s.h_count[q] = 0;
    }
    // ---- 5. records -> pairs, in windows of the protein-id space that fit shared memory -----------------------------
    for (uint32_t e = tid; e < E; e += T) {
        const uint32_t u = Cg[e];
        const uint32_t g0 = u == 0xffffffffu ? 0u : a.c.key_grp[u], g1 = u == 0xffffffffu ? 0u : a.c.key_grp[u + 1];
        Cg[e] = g0;  // current group
        B[e] = g1;   // end group
    }
    __syncthreads();
    uint32_t p_done = 0;
    uint32_t width = a.c.n_prot;  // protein-id span the last window covered
    while (true) {
        uint64_t mine = 0;
        for (uint32_t e = tid; e < E; e += T) mine += B[e] - Cg[e];
        const uint64_t remaining = block_sum64<T>(mine, s_w64);
        if (remaining == 0) break;
        uint32_t n_rec;
        if (remaining <= (uint64_t)RCAP) {
            for (uint32_t e = tid; e < E; e += T) D[e] = B[e];
            n_rec = (uint32_t)remaining;
        } else {
            // smallest target still to do; then the widest id window [pl, ph) whose records fit
            uint32_t pl_mine = 0xffffffffu;
            for (uint32_t e = tid; e < E; e += T) if (Cg[e] < B[e]) pl_mine = min(pl_mine, group_pid(a.c, Cg[e]));
            const uint32_t pl = block_min32<T>(pl_mine, s_warp);
            uint64_t span = min((uint64_t)width * 2, (uint64_t)a.c.n_prot - pl);
            if (span < 1) span = 1;
            while (true) {
                const uint64_t ph = (uint64_t)pl + span;  // exclusive
                uint64_t cnt = 0;
                for (uint32_t e = tid; e < E; e += T) {
                    uint32_t lo = Cg[e], hi = B[e];  // first group with pid >= ph
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if ((uint64_t)group_pid(a.c, mid) < ph) lo = mid + 1; else hi = mid;
                    }
                    D[e] = lo;
                    cnt += lo - Cg[e];
                }
                const uint64_t total = block_sum64<T>(cnt, s_w64);
                if (total <= (uint64_t)RCAP || span == 1) { n_rec = (uint32_t)total; break; }  // span 1: <= E <= RCAP records
                span = span / 2;
            }
            width = (uint32_t)min(span, (uint64_t)0xffffffffu);
        }
        __syncthreads();
        for (uint32_t e = tid; e < E; e += T) A[e] = D[e] - Cg[e];
        __syncthreads();
        block_scan<T>(A, E, s_warp);
        for (uint32_t r = tid; r < n_rec; r += T) {
            uint32_t lo = 0, hi = E - 1;  // last entry whose first record is <= r (entries without records share offsets)
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (A[mid] <= r) lo = mid; else hi = mid - 1;
            }
            const uint32_t g = Cg[lo] + (r - A[lo]);
            const uint32_t t0 = a.c.grp_start[g], t1 = a.c.grp_start[g + 1];
            s_key[r] = (a.c.loc[t0] & 0xffffffff00000000ull) | (uint64_t)(t1 - t0);
        }
        const uint32_t n_pad = pow2_at_least(n_rec);
        for (uint32_t i = n_rec + tid; i < n_pad; i += T) s_key[i] = ~0ull;
        __syncthreads();
        for (uint32_t e = tid; e < E; e += T) Cg[e] = D[e];  // this window's groups are consumed
        bitonic_sort<T, false>(s_key, nullptr, n_pad);
        for (uint32_t i = tid; i < n_rec; i += T) A[i] = (i == 0 || (s_key[i] >> 32) != (s_key[i - 1] >> 32)) ? 1u : 0u;
        __syncthreads();
        const uint32_t n_pairs = block_scan<T>(A, n_rec, s_warp);
        if (tid == 0) s_base = atomicAdd(reinterpret_cast<unsigned long long*>(a.s.totals + QT_CURSOR), (unsigned long long)n_pairs);
        __syncthreads();
        const uint64_t base = s_base;
        for (uint32_t i = tid; i < n_rec; i += T) {
            const uint32_t pid = (uint32_t)(s_key[i] >> 32);
            if (i != 0 && (uint32_t)(s_key[i - 1] >> 32) == pid) continue;  // not a head
            uint32_t j = i;
            uint64_t sum = 0;
            while (j < n_rec && (uint32_t)(s_key[j] >> 32) == pid) { sum += (uint32_t)s_key[j]; j++; }
            const uint32_t I = j - i;
            const double mean = (double)sum / (double)I;
            double var = 0.0;
            for (uint32_t t = i; t < j; t++) {
                const double d = (double)(uint32_t)s_key[t] - mean;
                var += d * d;
            }
            const uint32_t mid = i + I / 2;
            const double median = (I & 1u) ? (double)(uint32_t)s_key[mid]
                                           : ((double)(uint32_t)s_key[mid - 1] + (double)(uint32_t)s_key[mid]) * 0.5;
            const uint64_t slot = base + A[i];
            if (slot < a.s.stage_cap) {
                StagedPair sp;
                sp.qid = q; sp.rank = p_done + A[i]; sp.pid = pid; sp.isect = I; sp.sum_a = sum;
                sp.median = median; sp.stdev = sqrt(var / (double)I);
                a.s.stage[slot] = sp;
            }
        }
        p_done += n_pairs;
        __syncthreads();
    }
    if (tid == 0) a.s.p_count[q] = p_done;
}

// Exclusive scans over the queries of |Q|, pairs and hits; totals.  One CTA (the arrays are query-sized): every warp
// owns a contiguous segment of the queries and walks it 32 at a time with coalesced reads -- once to sum it, once (after
// the 32 segment sums are scanned) to write the offsets.  (Round 2 first scanned 1024 queries per iteration with every
// thread adding up all 32 warp totals of three 64-bit columns: 56 us for 10 000 queries, issue-bound on one SM.)
constexpr int QS_T = 1024;
__global__ void __launch_bounds__(QS_T) query_scan_kernel(QueryScratch s, uint32_t nq) {
    __shared__ uint64_t s_w[3][QS_T / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg = ((nq + 31) / 32 + 31) & ~31u;  // a multiple of 32: segments start on a 128-byte line
    const uint32_t q0 = min(warp * seg, nq), q1 = min(q0 + seg, nq);
    uint64_t sum[3] = {0, 0, 0};
#pragma unroll 4
    for (uint32_t q = q0 + lane; q < q1; q += 32) { sum[0] += s.e_count[q]; sum[1] += s.p_count[q]; sum[2] += s.h_count[q]; }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
        if (lane == 0) s_w[c][warp] = sum[c];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const uint64_t w = s_w[c][lane];
            uint64_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if ((int)lane >= o) wi += t;
            }
            s_w[c][lane] = wi - w;  // exclusive over the warps
            if (lane == 31) {
                if (c == 0) { s.sig_ptr[nq] = wi; s.totals[QT_ENTRIES] = wi; }
                if (c == 1) { s.pair_off[nq] = wi; s.totals[QT_PAIRS] = wi; }
                if (c == 2) { s.hit_off[nq] = wi; s.totals[QT_HITS] = wi; }
            }
        }
    }
    __syncthreads();
    uint64_t run[3] = {s_w[0][warp], s_w[1][warp], s_w[2][warp]};
    for (uint32_t base = q0; base < q1; base += 32) {
        const uint32_t q = base + lane;
        uint64_t v[3] = {0, 0, 0};
        if (q < q1) { v[0] = s.e_count[q]; v[1] = s.p_count[q]; v[2] = s.h_count[q]; }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            uint64_t incl = v[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            const uint64_t excl = run[c] + incl - v[c];
            if (q < q1) { if (c == 0) s.sig_ptr[q] = excl; else if (c == 1) s.pair_off[q] = excl; else s.hit_off[q] = excl; }
            run[c] += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

struct BlockPtrs {
    uint64_t* sig_ptr;
    uint32_t* u32[N_PAIR_U32];
    uint64_t* u64[N_PAIR_U64];
    double* score[N_SCORE_COLS];
    uint32_t* hit32[N_HIT_U32];
    uint64_t* hit_hash;
    uint64_t *q_mins, *q_abunds;
    uint64_t *pair_off, *hit_off;
    uint32_t* win_hoff;
};

BlockPtrs block_ptrs(const ResultLayout& L, void* block) {
    char* b = (char*)block;
    BlockPtrs p;
    p.sig_ptr = (uint64_t*)(b + L.off_sig_ptr);
    for (int i = 0; i < N_PAIR_U32; i++) p.u32[i] = (uint32_t*)(b + L.off_u32[i]);
    for (int i = 0; i < N_PAIR_U64; i++) p.u64[i] = (uint64_t*)(b + L.off_u64[i]);
    for (int i = 0; i < N_SCORE_COLS; i++) p.score[i] = (double*)(b + L.off_score[i]);
    for (int i = 0; i < N_HIT_U32; i++) p.hit32[i] = (uint32_t*)(b + L.off_hit32[i]);
    p.hit_hash = (uint64_t*)(b + L.off_hit_hash);
    p.q_mins = (uint64_t*)(b + L.off_q_mins);
    p.q_abunds = (uint64_t*)(b + L.off_q_abunds);
    p.pair_off = (uint64_t*)(b + L.off_pair_off);
    p.hit_off = (uint64_t*)(b + L.off_hit_off);
    p.win_hoff = (uint32_t*)(b + L.off_win_hoff);
    return p;
}

// The 12 fp64 columns of SURVEY App. A.6 from (|Q n T|, |Q|, |T|, sum of abundances, target total, median, deviation).
__device__ __forceinline__ void write_scores(const BlockPtrs& o, uint64_t j, uint32_t I, uint32_t qs, uint32_t ts, uint64_t sumA,
                                             uint64_t tw, double median, double stdev, uint32_t ksize) {
    const double dI = (double)I;
    const double cont = dI / (double)qs, cont_t = dI / (double)ts;
    const double inv = 1.0 / (double)(3u * ksize);
    const double qani = pow(cont, inv), mani = pow(cont_t, inv);
    o.score[SC_CONTAINMENT][j] = cont;
    o.score[SC_CONTAINMENT_TARGET][j] = cont_t;
    o.score[SC_MAX_CONTAINMENT][j] = fmax(cont, cont_t);
    o.score[SC_JACCARD][j] = dI / (double)((uint64_t)qs + ts - I);
    o.score[SC_QUERY_ANI][j] = qani;
    o.score[SC_MATCH_ANI][j] = mani;
    o.score[SC_AVERAGE_ANI][j] = (qani + mani) / 2.0;
    o.score[SC_MAX_ANI][j] = fmax(qani, mani);
    o.score[SC_AVERAGE_ABUND][j] = (double)sumA / dI;
    o.score[SC_MEDIAN_ABUND][j] = median;
    o.score[SC_STD_ABUND][j] = stdev;
    o.score[SC_F_WEIGHTED][j] = (double)sumA / (double)tw;
}

__global__ void finalize_pairs_kernel(CsrView c, QueryScratch s, uint64_t n_pairs, uint32_t ksize, uint32_t pid_base, BlockPtrs o) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const StagedPair sp = s.stage[i];
    const uint64_t j = s.pair_off[sp.qid] + sp.rank;
    const uint32_t qs = s.e_count[sp.qid], ts = c.t_size[sp.pid];
    const uint64_t tw = c.t_abund[sp.pid];
    o.u32[0][j] = sp.qid;
    o.u32[1][j] = sp.pid + pid_base;
    o.u32[2][j] = sp.isect;
    o.u32[3][j] = qs;
    o.u32[4][j] = ts;
    o.u64[0][j] = sp.sum_a;
    o.u64[1][j] = tw;
    write_scores(o, j, sp.isect, qs, ts, sp.sum_a, tw, sp.median, sp.stdev, ksize);
}

// last index i in [0, n) with a[i] <= v (a non-decreasing, a[0] <= v)
template <class T>
__device__ __forceinline__ uint32_t last_le(const T* __restrict__ a, uint32_t n, uint64_t v) {
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if ((uint64_t)a[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// One thread per hit: its query by a search over the queries' hit offsets, its window by a search over the window
// offsets of that query; the posting is row start + rest.  Order (query, qpos, target, tpos) by construction.
__global__ void expand_hits_kernel(CsrView c, QueryScratch s, const uint64_t* __restrict__ q_offs, uint32_t nq, uint32_t k,
                                   uint64_t n_hits, uint32_t pid_base, BlockPtrs o) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_hits) return;
    const uint32_t q = last_le(s.hit_off, nq, i);  // hit_off[q] <= i < hit_off[q + 1] (empty queries share offsets)
    const uint64_t ob = q_offs[q];
    const uint32_t W = (uint32_t)(q_offs[q + 1] - ob) - k + 1;
    const uint64_t local = i - s.hit_off[q];
    const uint32_t w = last_le(s.win_hoff + ob, W, local);
    const uint32_t u = s.win_key[ob + w];
    const uint64_t post = c.loc[c.grp_start[c.key_grp[u]] + (uint32_t)(local - s.win_hoff[ob + w])];
    o.hit32[0][i] = q;
    o.hit32[1][i] = (uint32_t)(post >> 32) + pid_base;
    o.hit32[2][i] = w;
    o.hit32[3][i] = (uint32_t)post;
    o.hit_hash[i] = c.keys[u];
}

// Sparse per-query sketches (at the query's residue offset) -> the compact concatenation behind sig_ptr.
__global__ void compact_sketches_kernel(QueryScratch s, const uint64_t* __restrict__ q_offs, uint32_t nq, uint64_t n_res, BlockPtrs o) {
    const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= n_res) return;
    const uint32_t q = last_le(q_offs, nq, g);  // offs[q] <= g (empty queries share offsets: the last one is the owner)
    const uint64_t e = g - q_offs[q];
    if (g >= q_offs[q + 1] || e >= s.e_count[q]) return;
    const uint64_t d = s.sig_ptr[q] + e;
    o.q_mins[d] = s.ent_hash[g];
    o.q_abunds[d] = s.ent_abund[g];
}

template <int QCAP, int RCAP, int T>
cudaError_t launch_query_variant(const QueryKernelArgs& a, const Lut256& lut, cudaStream_t st) {
    constexpr size_t smem = query_smem_bytes<QCAP, RCAP>();
    cudaError_t e = cudaFuncSetAttribute(query_kernel<QCAP, RCAP, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    query_kernel<QCAP, RCAP, T><<<a.nq, T, smem, st>>>(a, lut);
    return cudaGetLastError();
}

}  // namespace

ResultLayout result_layout(uint64_t nq, uint64_t n_pairs, uint64_t n_hits, uint64_t n_entries, bool hits, bool sketches,
                           bool wire, uint64_t n_res) {
    ResultLayout L;
    L.nq = nq; L.n_pairs = n_pairs; L.n_hits = hits ? n_hits : 0; L.n_entries = sketches ? n_entries : 0;
    L.hits = hits ? 1 : 0; L.sketches = sketches ? 1 : 0; L.wire = wire ? 1 : 0; L.n_res = n_res;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes ? bytes : 1); return o; };
    L.off_sig_ptr = take((nq + 1) * 8);
    for (int i = 0; i < N_PAIR_U32; i++) L.off_u32[i] = take(n_pairs * 4);
    for (int i = 0; i < N_PAIR_U64; i++) L.off_u64[i] = take(n_pairs * 8);
    for (int i = 0; i < N_SCORE_COLS; i++) L.off_score[i] = take(n_pairs * 8);
    for (int i = 0; i < N_HIT_U32; i++) L.off_hit32[i] = take(L.n_hits * 4);
    L.off_hit_hash = take(L.n_hits * 8);
    L.off_q_mins = take(L.n_entries * 8);
    L.off_q_abunds = take(L.n_entries * 8);
    if (wire) {
        L.off_pair_off = take((nq + 1) * 8);
        L.off_hit_off = take(hits ? (nq + 1) * 8 : 0);
        L.off_win_hoff = take(hits ? n_res * 4 : 0);
    }
    L.bytes = off;
    return L;
}

cudaError_t launch_query_phase1(const QueryBatchView& q, const CsrView& csr, uint32_t k, int moltype, uint64_t max_hash,
                                bool want_hits, const QueryScratch& s, cudaStream_t stream, uint64_t* n_launches) {
#define KS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
    KS_TRY(cudaMemsetAsync(s.totals, 0, QT_WORDS * 8, stream));
    if (q.nq) {
        Lut256 lut;
        fill_lut(moltype, &lut);
        QueryKernelArgs a;
        a.res = q.res; a.offs = q.offs; a.nq = q.nq; a.k = k; a.max_hash = max_hash; a.c = csr; a.s = s;
        a.want_hits = want_hits ? 1 : 0;
        a.w_lo = 0; a.w_hi = QK_SMALL_WINDOWS;
        KS_TRY((launch_query_variant<(int)QK_SMALL_WINDOWS, 1024, 128>(a, lut, stream)));
        if (n_launches) *n_launches += 1;
        if (q.max_windows > QK_SMALL_WINDOWS) {
            a.w_lo = QK_SMALL_WINDOWS; a.w_hi = QK_MAX_WINDOWS;
            KS_TRY((launch_query_variant<(int)QK_MAX_WINDOWS, (int)QK_MAX_WINDOWS, 512>(a, lut, stream)));
            if (n_launches) *n_launches += 1;
        }
    }
    query_scan_kernel<<<1, QS_T, 0, stream>>>(s, q.nq);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_query_phase2(const QueryBatchView& q, const CsrView& csr, uint32_t k, uint32_t pid_base,
                                const QueryScratch& s, const ResultLayout& L, void* block, cudaStream_t stream,
                                uint64_t* n_launches) {
    const BlockPtrs o = block_ptrs(L, block);
    KS_TRY(cudaMemcpyAsync(o.sig_ptr, s.sig_ptr, (L.nq + 1) * 8, cudaMemcpyDeviceToDevice, stream));
    if (L.n_pairs) {
        finalize_pairs_kernel<<<blocks_for(L.n_pairs), TB, 0, stream>>>(csr, s, L.n_pairs, k, pid_base, o);
        KS_TRY(cudaGetLastError());
        if (n_launches) *n_launches += 1;
    }
    if (L.hits && L.n_hits) {
        expand_hits_kernel<<<blocks_for(L.n_hits), TB, 0, stream>>>(csr, s, q.offs, q.nq, k, L.n_hits, pid_base, o);
        KS_TRY(cudaGetLastError());
        if (n_launches) *n_launches += 1;
    }
    if (L.sketches && L.n_entries) {
        compact_sketches_kernel<<<blocks_for(q.n_res), TB, 0, stream>>>(s, q.offs, q.nq, q.n_res, o);
        KS_TRY(cudaGetLastError());
        if (n_launches) *n_launches += 1;
    }
    if (L.wire) {
        KS_TRY(cudaMemcpyAsync(o.pair_off, s.pair_off, (L.nq + 1) * 8, cudaMemcpyDeviceToDevice, stream));
        if (L.hits) {
            KS_TRY(cudaMemcpyAsync(o.hit_off, s.hit_off, (L.nq + 1) * 8, cudaMemcpyDeviceToDevice, stream));
            if (L.n_res) KS_TRY(cudaMemcpyAsync(o.win_hoff, s.win_hoff, L.n_res * 4, cudaMemcpyDeviceToDevice, stream));
        }
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------------------------
// multi-GPU merge on rank 0
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct MergeDev {
    int n_shards;
    uint32_t nq, k;
    const uint64_t* q_offs;
    BlockPtrs in[MAX_SHARDS];
    uint64_t n_pairs[MAX_SHARDS], n_hits[MAX_SHARDS];
    BlockPtrs out;
};

// grid.y = shard; one thread per pair of that shard
__global__ void merge_pairs_kernel(const MergeDev* __restrict__ mp) {
    const MergeDev& m = *mp;
    const int s = blockIdx.y;
    const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= m.n_pairs[s]) return;
    const BlockPtrs& in = m.in[s];
    const uint32_t q = in.u32[0][j];
    uint64_t dst = j - in.pair_off[q];
    for (int t = 0; t < m.n_shards; t++) {
        const uint64_t a = m.in[t].pair_off[q];
        dst += a;
        if (t < s) dst += m.in[t].pair_off[q + 1] - a;
    }
#pragma unroll
    for (int i = 0; i < N_PAIR_U32; i++) m.out.u32[i][dst] = in.u32[i][j];
#pragma unroll
    for (int i = 0; i < N_PAIR_U64; i++) m.out.u64[i][dst] = in.u64[i][j];
#pragma unroll
    for (int i = 0; i < N_SCORE_COLS; i++) m.out.score[i][dst] = in.score[i][j];
}

// hits of shard t before window (q, w) in the shard's own order
__device__ __forceinline__ uint64_t shard_hits_before(const BlockPtrs& in, uint64_t ob, uint32_t q, uint32_t w, uint32_t W) {
    return w < W ? in.hit_off[q] + in.win_hoff[ob + w] : in.hit_off[q + 1];
}

__global__ void merge_hits_kernel(const MergeDev* __restrict__ mp) {
    const MergeDev& m = *mp;
    const int s = blockIdx.y;
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m.n_hits[s]) return;
    const BlockPtrs& in = m.in[s];
    const uint32_t q = in.hit32[0][i], w = in.hit32[2][i];
    const uint64_t ob = m.q_offs[q];
    const uint32_t W = (uint32_t)(m.q_offs[q + 1] - ob) - m.k + 1;
    uint64_t dst = i - shard_hits_before(in, ob, q, w, W);
    for (int t = 0; t < m.n_shards; t++) {
        const uint64_t a = shard_hits_before(m.in[t], ob, q, w, W);
        dst += a;
        if (t < s) dst += shard_hits_before(m.in[t], ob, q, w + 1, W) - a;
    }
#pragma unroll
    for (int c = 0; c < N_HIT_U32; c++) m.out.hit32[c][dst] = in.hit32[c][i];
    m.out.hit_hash[dst] = in.hit_hash[i];
}

}  // namespace

cudaError_t launch_merge(const MergeArgs& m, cudaStream_t stream, uint64_t* n_launches) {
    if (m.n_shards < 1 || m.n_shards > MAX_SHARDS) return cudaErrorInvalidValue;
    MergeDev h;
    h.n_shards = m.n_shards; h.nq = m.nq; h.k = m.k; h.q_offs = m.q_offs;
    uint64_t max_pairs = 0, max_hits = 0;
    for (int s = 0; s < m.n_shards; s++) {
        h.in[s] = block_ptrs(m.shard[s].layout, const_cast<void*>(m.shard[s].block));
        h.n_pairs[s] = m.shard[s].layout.n_pairs;
        h.n_hits[s] = m.shard[s].layout.hits ? m.shard[s].layout.n_hits : 0;
        max_pairs = std::max(max_pairs, h.n_pairs[s]);
        max_hits = std::max(max_hits, h.n_hits[s]);
    }
    h.out = block_ptrs(m.out_layout, m.out_block);
    MergeDev* d = nullptr;  // the descriptor (17 sets of column pointers) is too large for kernel parameters
    KS_TRY(cudaMallocAsync(&d, sizeof(MergeDev), stream));
    KS_TRY(cudaMemcpyAsync(d, &h, sizeof(MergeDev), cudaMemcpyHostToDevice, stream));
    // (the copy reads pageable host memory: it is staged by the runtime before the call returns)
    const BlockPtrs first = h.in[0];
    KS_TRY(cudaMemcpyAsync(h.out.sig_ptr, first.sig_ptr, (m.nq + 1) * 8, cudaMemcpyDeviceToDevice, stream));
    if (m.out_layout.sketches && m.out_layout.n_entries) {
        KS_TRY(cudaMemcpyAsync(h.out.q_mins, first.q_mins, m.out_layout.n_entries * 8, cudaMemcpyDeviceToDevice, stream));
        KS_TRY(cudaMemcpyAsync(h.out.q_abunds, first.q_abunds, m.out_layout.n_entries * 8, cudaMemcpyDeviceToDevice, stream));
    }
    if (max_pairs) {
        merge_pairs_kernel<<<dim3(blocks_for(max_pairs), m.n_shards), TB, 0, stream>>>(d);
        KS_TRY(cudaGetLastError());
        if (n_launches) *n_launches += 1;
    }
    if (m.out_layout.hits && max_hits) {
        merge_hits_kernel<<<dim3(blocks_for(max_hits), m.n_shards), TB, 0, stream>>>(d);
        KS_TRY(cudaGetLastError());
        if (n_launches) *n_launches += 1;
    }
    return cudaFreeAsync(d, stream);
#undef KS_TRY
}

namespace {
__device__ __forceinline__ void id_sum_key(const CsrView& v, uint32_t u, unsigned long long* sums) {
    const uint64_t h = v.keys[u];
    for (uint32_t g = v.key_grp[u]; g < v.key_grp[u + 1]; g++) atomicAdd(sums + (uint32_t)(v.loc[v.grp_start[g]] >> 32), (unsigned long long)h);
}
__global__ void id_sums_compact_kernel(CsrView v, uint64_t n_keys, unsigned long long* sums) {
    for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < n_keys; u += (uint64_t)gridDim.x * blockDim.x)
        id_sum_key(v, (uint32_t)u, sums);
}
__global__ void id_sums_seg_kernel(CsrView v, unsigned long long* sums) {
    const uint32_t b = blockIdx.x;
    const uint32_t kb = v.seg_start[b] + b, tk = (uint32_t)v.seg_counts[b];
    for (uint32_t i = threadIdx.x; i < tk; i += blockDim.x) id_sum_key(v, kb + i, sums);
}
}  // namespace

cudaError_t launch_id_sums(const CsrView& v, uint64_t n_keys, unsigned long long* sums, cudaStream_t stream) {
    if (n_keys == 0) return cudaSuccess;
    if (v.dir_sub != DIR_SUB_COMPACT) id_sums_seg_kernel<<<v.seg_nb, 128, 0, stream>>>(v, sums);
    else id_sums_compact_kernel<<<148 * 8, 256, 0, stream>>>(v, n_keys, sums);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// library-sorted path (round 1): grouping for the export calls, and the search of batches with very long queries
// ------------------------------------------------------------------------------------------------------------------
namespace {

void exclusive_scan(Arena& tmp, const uint64_t* in, uint64_t* out, uint64_t n, uint64_t* n_launches) {
    size_t bytes = 0;
    KS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n, tmp.stream()));
    void* t = tmp.alloc<char>(bytes);
    KS_CUDA(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, (int64_t)n, tmp.stream()));
    tmp.release(t);
    if (n_launches) *n_launches += 2;
}

template <class KeyT, class ValT>
void sort_pairs(Arena& tmp, KeyT*& keys, ValT*& vals, uint64_t n, int begin_bit, int end_bit, uint64_t* n_launches) {
    if (n == 0) return;
    KeyT* kb = tmp.alloc<KeyT>(n);
    ValT* vb = tmp.alloc<ValT>(n);
    cub::DoubleBuffer<KeyT> k(keys, kb);
    cub::DoubleBuffer<ValT> v(vals, vb);
    size_t bytes = 0;
    KS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int64_t)n, begin_bit, end_bit, tmp.stream()));
    void* t = tmp.alloc<char>(bytes);
    KS_CUDA(cub::DeviceRadixSort::SortPairs(t, bytes, k, v, (int64_t)n, begin_bit, end_bit, tmp.stream()));
    tmp.release(t);
    // whichever buffer holds the result becomes the caller's; the other is scratch
    if (k.Current() != keys) { tmp.release(keys); keys = kb; } else { tmp.release(kb); }
    if (v.Current() != vals) { tmp.release(vals); vals = vb; } else { tmp.release(vb); }
    if (n_launches) *n_launches += 2 + (end_bit - begin_bit + 7) / 8;
}

uint64_t read_u64(const uint64_t* d, cudaStream_t st) {
    uint64_t v = 0;
    KS_CUDA(cudaMemcpyAsync(&v, d, 8, cudaMemcpyDeviceToHost, st));
    KS_CUDA(cudaStreamSynchronize(st));
    return v;
}

// ---- grouping tuples into per-owner sketches ------------------------------------------------
__global__ void flag_entries_kernel(const uint64_t* __restrict__ h, const uint64_t* __restrict__ loc, uint64_t n,
                                    uint64_t* __restrict__ flags) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { flags[i] = 0; return; }
    flags[i] = (i == 0 || h[i] != h[i - 1] || (loc[i] >> 32) != (loc[i - 1] >> 32)) ? 1 : 0;
}

__global__ void scatter_entries_kernel(const uint64_t* __restrict__ h, const uint64_t* __restrict__ loc, uint64_t n,
                                       const uint64_t* __restrict__ flags, const uint64_t* __restrict__ pos,
                                       uint64_t* __restrict__ ent_hash, uint32_t* __restrict__ ent_owner,
                                       uint32_t* __restrict__ ent_first) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { ent_first[pos[n]] = (uint32_t)n; return; }
    if (flags[i]) {
        uint64_t e = pos[i];
        ent_hash[e] = h[i];
        ent_owner[e] = (uint32_t)(loc[i] >> 32);
        ent_first[e] = (uint32_t)i;
    }
}

__global__ void owner_ptr_kernel(const uint32_t* __restrict__ ent_owner, uint64_t E, uint32_t n_owner,
                                 uint64_t* __restrict__ sig_ptr) {
    uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o > n_owner) return;
    uint64_t lo = 0, hi = E;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (ent_owner[mid] < o) lo = mid + 1; else hi = mid;
    }
    sig_ptr[o] = lo;
}

__global__ void lookup_entries_kernel(CsrView c, const uint64_t* __restrict__ ent_hash, uint64_t E,
                                      uint32_t* __restrict__ ent_key, uint64_t* __restrict__ ent_ngrp) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e > E) return;
    if (e == E) { ent_ngrp[e] = 0; return; }
    uint32_t u = find_key(c, ent_hash[e]);
    ent_key[e] = u;
    ent_ngrp[e] = u == 0xffffffffu ? 0 : (uint64_t)(c.key_grp[u + 1] - c.key_grp[u]);
}

// One record per (query entry, target group): key = (query << 32) | protein, cnt = abundance of the hash in
// the target's sketch.  Output slot r is mapped back to its source entry by a search in the scanned counts.
__global__ void expand_groups_kernel(CsrView c, const uint64_t* __restrict__ ent_goff, uint64_t E, uint64_t R,
                                     const uint32_t* __restrict__ ent_key, const uint32_t* __restrict__ ent_owner,
                                     uint64_t* __restrict__ rec_key, uint32_t* __restrict__ rec_cnt) {
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= R) return;
    uint64_t lo = 0, hi = E;  // last e with ent_goff[e] <= r
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (ent_goff[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const uint32_t g = c.key_grp[ent_key[lo]] + (uint32_t)(r - ent_goff[lo]);
    const uint32_t t0 = c.grp_start[g], t1 = c.grp_start[g + 1];
    rec_key[r] = ((uint64_t)ent_owner[lo] << 32) | (c.loc[t0] >> 32);
    rec_cnt[r] = t1 - t0;
}

__global__ void flag_pairs_kernel(const uint64_t* __restrict__ rec_key, uint64_t R, uint64_t* __restrict__ flags) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i > R) return;
    flags[i] = (i < R && (i == 0 || rec_key[i] != rec_key[i - 1])) ? 1 : 0;
}

__global__ void scatter_pairs_kernel(const uint64_t* __restrict__ flags, const uint64_t* __restrict__ pos, uint64_t R,
                                     uint64_t* __restrict__ pair_start) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i > R) return;
    if (i == R) { pair_start[pos[R]] = R; return; }
    if (flags[i]) pair_start[pos[i]] = i;
}

// One thread per scored pair; its records are contiguous, ordered by abundance (for the median).
__global__ void legacy_score_kernel(CsrView c, const uint64_t* __restrict__ pair_start, uint64_t n_pairs,
                                    const uint64_t* __restrict__ rec_key, const uint32_t* __restrict__ rec_cnt,
                                    const uint64_t* __restrict__ q_sig_ptr, uint32_t ksize, uint32_t pid_base, BlockPtrs o) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n_pairs) return;
    const uint64_t a = pair_start[j], b = pair_start[j + 1];
    const uint64_t key = rec_key[a];
    const uint32_t qid = (uint32_t)(key >> 32), pid = (uint32_t)key;
    const uint64_t I = b - a;
    uint64_t sumA = 0;
    for (uint64_t i = a; i < b; i++) sumA += rec_cnt[i];
    const double dI = (double)I;
    const double mean = (double)sumA / dI;
    double var = 0.0;
    for (uint64_t i = a; i < b; i++) {
        double d = (double)rec_cnt[i] - mean;
        var += d * d;
    }
    const uint64_t mid = a + I / 2;
    const double median = (I & 1) ? (double)rec_cnt[mid] : ((double)rec_cnt[mid - 1] + (double)rec_cnt[mid]) * 0.5;
    const uint32_t qs = (uint32_t)(q_sig_ptr[qid + 1] - q_sig_ptr[qid]);
    const uint32_t ts = c.t_size[pid];
    const uint64_t tw = c.t_abund[pid];
    o.u32[0][j] = qid;
    o.u32[1][j] = pid + pid_base;
    o.u32[2][j] = (uint32_t)I;
    o.u32[3][j] = qs;
    o.u32[4][j] = ts;
    o.u64[0][j] = sumA;
    o.u64[1][j] = tw;
    write_scores(o, j, (uint32_t)I, qs, ts, sumA, tw, median, sqrt(var / dI), ksize);
}

__global__ void lookup_tuples_kernel(CsrView c, const uint64_t* __restrict__ q_hash, uint64_t n,
                                     uint32_t* __restrict__ row0, uint64_t* __restrict__ row_len) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t > n) return;
    if (t == n) { row_len[t] = 0; return; }
    uint32_t u = find_key(c, q_hash[t]);
    if (u == 0xffffffffu) { row0[t] = 0; row_len[t] = 0; return; }
    const uint32_t a = c.grp_start[c.key_grp[u]], b = c.grp_start[c.key_grp[u + 1]];
    row0[t] = a;
    row_len[t] = b - a;
}

__global__ void legacy_expand_hits_kernel(CsrView c, const uint64_t* __restrict__ q_hash, const uint64_t* __restrict__ q_loc,
                                          uint64_t n, const uint32_t* __restrict__ row0, const uint64_t* __restrict__ off,
                                          uint64_t H, uint32_t pid_base, BlockPtrs o) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= H) return;
    uint64_t lo = 0, hi = n;  // last t with off[t] <= i
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const uint64_t post = c.loc[row0[lo] + (i - off[lo])];
    const uint64_t ql = q_loc[lo];
    o.hit32[0][i] = (uint32_t)(ql >> 32);
    o.hit32[1][i] = (uint32_t)(post >> 32) + pid_base;
    o.hit32[2][i] = (uint32_t)ql;
    o.hit32[3][i] = (uint32_t)post;
    o.hit_hash[i] = q_hash[lo];
}

__global__ void abund_u64_kernel(const uint32_t* __restrict__ ent_first, uint64_t E, uint64_t* __restrict__ out) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e < E) out[e] = ent_first[e + 1] - ent_first[e];
}

}  // namespace

void group_by_owner(Arena& keep, Arena& tmp, const uint64_t* hash, const uint64_t* loc, uint64_t n, uint32_t n_owner,
                    int hash_end_bit, Grouped* out, uint64_t* n_launches) {
    cudaStream_t st = keep.stream();
    Grouped g;
    g.n = n;
    // copies, then: stable sort by hash, stable sort by owner  ->  (owner, hash, pos)
    uint64_t* sh = tmp.alloc<uint64_t>(n);
    uint64_t* sl = tmp.alloc<uint64_t>(n);
    if (n) {
        KS_CUDA(cudaMemcpyAsync(sh, hash, n * 8, cudaMemcpyDeviceToDevice, st));
        KS_CUDA(cudaMemcpyAsync(sl, loc, n * 8, cudaMemcpyDeviceToDevice, st));
    }
    sort_pairs(tmp, sh, sl, n, 0, hash_end_bit, n_launches);
    sort_pairs(tmp, sl, sh, n, 32, 32 + bits_for(n_owner ? n_owner - 1 : 0), n_launches);
    uint64_t* flags = tmp.alloc<uint64_t>(n + 1);
    uint64_t* pos = tmp.alloc<uint64_t>(n + 1);
    flag_entries_kernel<<<blocks_for(n + 1), TB, 0, st>>>(sh, sl, n, flags);
    KS_CUDA(cudaGetLastError());
    exclusive_scan(tmp, flags, pos, n + 1, n_launches);
    const uint64_t E = read_u64(pos + n, st);
    g.n_entries = E;
    g.ent_hash = keep.alloc<uint64_t>(E);
    g.ent_owner = keep.alloc<uint32_t>(E);
    g.ent_first = keep.alloc<uint32_t>(E + 1);
    g.sig_ptr = keep.alloc<uint64_t>((uint64_t)n_owner + 1);
    scatter_entries_kernel<<<blocks_for(n + 1), TB, 0, st>>>(sh, sl, n, flags, pos, g.ent_hash, g.ent_owner, g.ent_first);
    KS_CUDA(cudaGetLastError());
    owner_ptr_kernel<<<blocks_for((uint64_t)n_owner + 1), TB, 0, st>>>(g.ent_owner, E, n_owner, g.sig_ptr);
    KS_CUDA(cudaGetLastError());
    if (n_launches) *n_launches += 3;
    tmp.release(flags);
    tmp.release(pos);
    tmp.move_to(keep, sh);
    tmp.move_to(keep, sl);
    g.s_hash = sh;
    g.s_loc = sl;
    *out = g;
}

void search_device_legacy(Arena& tmp, const CsrView& csr, const uint64_t* q_hash, const uint64_t* q_loc, uint64_t nqt,
                          uint32_t n_queries, uint32_t ksize, int hash_end_bit, bool want_hits, bool want_sketches,
                          uint32_t pid_base, Arena& block_owner, void** block_out, ResultLayout* layout_out,
                          uint64_t* n_launches) {
    cudaStream_t st = tmp.stream();
    Grouped qs;
    group_by_owner(tmp, tmp, q_hash, q_loc, nqt, n_queries, hash_end_bit, &qs, n_launches);
    const uint64_t E = qs.n_entries;

    // 1. one lookup per distinct (query, hash); count target groups behind each
    uint32_t* ent_key = tmp.alloc<uint32_t>(E);
    uint64_t* ent_ngrp = tmp.alloc<uint64_t>(E + 1);
    uint64_t* ent_goff = tmp.alloc<uint64_t>(E + 1);
    lookup_entries_kernel<<<blocks_for(E + 1), TB, 0, st>>>(csr, qs.ent_hash, E, ent_key, ent_ngrp);
    KS_CUDA(cudaGetLastError());
    exclusive_scan(tmp, ent_ngrp, ent_goff, E + 1, n_launches);
    const uint64_t R = read_u64(ent_goff + E, st);

    // 2. expand to (query, protein, abundance) records; order by (query, protein, abundance)
    uint64_t* rec_key = tmp.alloc<uint64_t>(R);
    uint32_t* rec_cnt = tmp.alloc<uint32_t>(R);
    if (R) {
        expand_groups_kernel<<<blocks_for(R), TB, 0, st>>>(csr, ent_goff, E, R, ent_key, qs.ent_owner, rec_key, rec_cnt);
        KS_CUDA(cudaGetLastError());
        sort_pairs(tmp, rec_cnt, rec_key, R, 0, 32, n_launches);
        sort_pairs(tmp, rec_key, rec_cnt, R, 0, 32 + bits_for(n_queries ? n_queries - 1 : 0), n_launches);
    }
    if (n_launches) *n_launches += 2;

    // 3. pair boundaries
    uint64_t* flags = tmp.alloc<uint64_t>(R + 1);
    uint64_t* pos = tmp.alloc<uint64_t>(R + 1);
    flag_pairs_kernel<<<blocks_for(R + 1), TB, 0, st>>>(rec_key, R, flags);
    KS_CUDA(cudaGetLastError());
    exclusive_scan(tmp, flags, pos, R + 1, n_launches);
    const uint64_t NP = read_u64(pos + R, st);
    uint64_t* pair_start = tmp.alloc<uint64_t>(NP + 1);
    scatter_pairs_kernel<<<blocks_for(R + 1), TB, 0, st>>>(flags, pos, R, pair_start);
    KS_CUDA(cudaGetLastError());

    // 4. hit rows (sizes first: the block is allocated once)
    uint32_t* row0 = nullptr;
    uint64_t* off = nullptr;
    uint64_t H = 0;
    if (want_hits) {
        row0 = tmp.alloc<uint32_t>(nqt);
        uint64_t* row_len = tmp.alloc<uint64_t>(nqt + 1);
        off = tmp.alloc<uint64_t>(nqt + 1);
        lookup_tuples_kernel<<<blocks_for(nqt + 1), TB, 0, st>>>(csr, q_hash, nqt, row0, row_len);
        KS_CUDA(cudaGetLastError());
        exclusive_scan(tmp, row_len, off, nqt + 1, n_launches);
        H = read_u64(off + nqt, st);
    }
    const ResultLayout L = result_layout(n_queries, NP, H, E, want_hits, want_sketches);
    void* block = block_owner.alloc<char>(L.bytes);
    const BlockPtrs o = block_ptrs(L, block);
    KS_CUDA(cudaMemcpyAsync(o.sig_ptr, qs.sig_ptr, ((uint64_t)n_queries + 1) * 8, cudaMemcpyDeviceToDevice, st));
    if (NP) {
        legacy_score_kernel<<<blocks_for(NP), TB, 0, st>>>(csr, pair_start, NP, rec_key, rec_cnt, qs.sig_ptr, ksize, pid_base, o);
        KS_CUDA(cudaGetLastError());
    }
    if (want_hits && H) {
        legacy_expand_hits_kernel<<<blocks_for(H), TB, 0, st>>>(csr, q_hash, q_loc, nqt, row0, off, H, pid_base, o);
        KS_CUDA(cudaGetLastError());
    }
    if (want_sketches && E) {
        KS_CUDA(cudaMemcpyAsync(o.q_mins, qs.ent_hash, E * 8, cudaMemcpyDeviceToDevice, st));
        abund_u64_kernel<<<blocks_for(E), TB, 0, st>>>(qs.ent_first, E, o.q_abunds);
        KS_CUDA(cudaGetLastError());
    }
    if (n_launches) *n_launches += 6;
    *block_out = block;
    *layout_out = L;
}

}  // namespace ks
