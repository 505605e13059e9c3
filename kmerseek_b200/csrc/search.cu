// Query kernels for sm_100a: batched lookup of query sketches in the CSR index, per-(query, target)
// aggregation and fp64 scores, and the hit list (query pos x target pos).
//
// Replaces sourmash_plugin_branchwater.do_manysearch as called at src/python/kmerseek/search.py:125-141
// (all-pairs sorted-list intersections) and the polars join on (encoded, hashval) at search.py:204-213.
// Score formulae: SURVEY.md Appendix A.6.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "search.cuh"

namespace ks {

namespace {

constexpr int TB = 256;
inline unsigned blocks_for(uint64_t n) { return (unsigned)((n + TB - 1) / TB); }

inline int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b)) b++;
    return b;
}

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

// ---- lookups ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_key(const CsrView& c, uint64_t h) {
    const uint64_t b = h >> c.dir_shift;
    if (b >= (1ull << c.dir_bits)) return 0xffffffffu;
    uint32_t lo = c.dir[b], hi = c.dir[b + 1];
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (c.keys[mid] < h) lo = mid + 1; else hi = mid;
    }
    return (lo < c.dir[b + 1] && c.keys[lo] == h) ? lo : 0xffffffffu;
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

struct ScoreOut {
    uint32_t *pair_qid, *pair_pid, *intersect, *q_size, *t_size;
    uint64_t *nwf, *twh;
    double* s[N_SCORE_COLS];
};

// One thread per scored pair; its records are contiguous, ordered by abundance (for the median).
__global__ void score_kernel(CsrView c, const uint64_t* __restrict__ pair_start, uint64_t n_pairs,
                             const uint64_t* __restrict__ rec_key, const uint32_t* __restrict__ rec_cnt,
                             const uint64_t* __restrict__ q_sig_ptr, uint32_t ksize, ScoreOut o) {
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
    const double cont = dI / (double)qs, cont_t = dI / (double)ts;
    const double inv = 1.0 / (double)(3u * ksize);
    const double qani = pow(cont, inv), mani = pow(cont_t, inv);
    o.pair_qid[j] = qid;
    o.pair_pid[j] = pid;
    o.intersect[j] = (uint32_t)I;
    o.q_size[j] = qs;
    o.t_size[j] = ts;
    o.nwf[j] = sumA;
    o.twh[j] = tw;
    o.s[SC_CONTAINMENT][j] = cont;
    o.s[SC_CONTAINMENT_TARGET][j] = cont_t;
    o.s[SC_MAX_CONTAINMENT][j] = fmax(cont, cont_t);
    o.s[SC_JACCARD][j] = dI / (double)((uint64_t)qs + ts - I);
    o.s[SC_QUERY_ANI][j] = qani;
    o.s[SC_MATCH_ANI][j] = mani;
    o.s[SC_AVERAGE_ANI][j] = (qani + mani) / 2.0;
    o.s[SC_MAX_ANI][j] = fmax(qani, mani);
    o.s[SC_AVERAGE_ABUND][j] = mean;
    o.s[SC_MEDIAN_ABUND][j] = median;
    o.s[SC_STD_ABUND][j] = sqrt(var / dI);
    o.s[SC_F_WEIGHTED][j] = (double)sumA / (double)tw;
}

// ---- hit list -----------------------------------------------------------------------------------
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

__global__ void expand_hits_kernel(CsrView c, const uint64_t* __restrict__ q_hash, const uint64_t* __restrict__ q_loc,
                                   uint64_t n, const uint32_t* __restrict__ row0, const uint64_t* __restrict__ off,
                                   uint64_t H, uint32_t* __restrict__ hit_qid, uint32_t* __restrict__ hit_pid,
                                   uint64_t* __restrict__ hit_hash, uint32_t* __restrict__ hit_qpos,
                                   uint32_t* __restrict__ hit_tpos) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= H) return;
    uint64_t lo = 0, hi = n;  // last t with off[t] <= i
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const uint64_t post = c.loc[row0[lo] + (i - off[lo])];
    const uint64_t ql = q_loc[lo];
    hit_qid[i] = (uint32_t)(ql >> 32);
    hit_qpos[i] = (uint32_t)ql;
    hit_hash[i] = q_hash[lo];
    hit_pid[i] = (uint32_t)(post >> 32);
    hit_tpos[i] = (uint32_t)post;
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

void search_device(Arena& keep, Arena& tmp, const CsrView& csr, const uint64_t* q_hash, const uint64_t* q_loc,
                   uint64_t nqt, uint32_t n_queries, uint32_t ksize, int hash_end_bit, bool want_hits,
                   Grouped* qs, SearchDevice* out, uint64_t* n_launches) {
    cudaStream_t st = keep.stream();
    group_by_owner(keep, tmp, q_hash, q_loc, nqt, n_queries, hash_end_bit, qs, n_launches);
    const uint64_t E = qs->n_entries;
    SearchDevice o;

    // 1. one lookup per distinct (query, hash); count target groups behind each
    uint32_t* ent_key = tmp.alloc<uint32_t>(E);
    uint64_t* ent_ngrp = tmp.alloc<uint64_t>(E + 1);
    uint64_t* ent_goff = tmp.alloc<uint64_t>(E + 1);
    lookup_entries_kernel<<<blocks_for(E + 1), TB, 0, st>>>(csr, qs->ent_hash, E, ent_key, ent_ngrp);
    KS_CUDA(cudaGetLastError());
    exclusive_scan(tmp, ent_ngrp, ent_goff, E + 1, n_launches);
    const uint64_t R = read_u64(ent_goff + E, st);

    // 2. expand to (query, protein, abundance) records; order by (query, protein, abundance)
    uint64_t* rec_key = tmp.alloc<uint64_t>(R);
    uint32_t* rec_cnt = tmp.alloc<uint32_t>(R);
    if (R) {
        expand_groups_kernel<<<blocks_for(R), TB, 0, st>>>(csr, ent_goff, E, R, ent_key, qs->ent_owner, rec_key, rec_cnt);
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

    // 4. scores
    o.n_pairs = NP;
    o.pair_qid = keep.alloc<uint32_t>(NP);
    o.pair_pid = keep.alloc<uint32_t>(NP);
    o.intersect = keep.alloc<uint32_t>(NP);
    o.q_size = keep.alloc<uint32_t>(NP);
    o.t_size = keep.alloc<uint32_t>(NP);
    o.n_weighted_found = keep.alloc<uint64_t>(NP);
    o.total_weighted = keep.alloc<uint64_t>(NP);
    ScoreOut so;
    so.pair_qid = o.pair_qid; so.pair_pid = o.pair_pid; so.intersect = o.intersect; so.q_size = o.q_size;
    so.t_size = o.t_size; so.nwf = o.n_weighted_found; so.twh = o.total_weighted;
    for (int i = 0; i < N_SCORE_COLS; i++) so.s[i] = o.score[i] = keep.alloc<double>(NP);
    if (NP) {
        score_kernel<<<blocks_for(NP), TB, 0, st>>>(csr, pair_start, NP, rec_key, rec_cnt, qs->sig_ptr, ksize, so);
        KS_CUDA(cudaGetLastError());
    }
    if (n_launches) *n_launches += 3;

    // 5. hit list: every (query occurrence, posting) of a shared hash, in (query, qpos, protein, tpos) order
    if (want_hits) {
        uint32_t* row0 = tmp.alloc<uint32_t>(nqt);
        uint64_t* row_len = tmp.alloc<uint64_t>(nqt + 1);
        uint64_t* off = tmp.alloc<uint64_t>(nqt + 1);
        lookup_tuples_kernel<<<blocks_for(nqt + 1), TB, 0, st>>>(csr, q_hash, nqt, row0, row_len);
        KS_CUDA(cudaGetLastError());
        exclusive_scan(tmp, row_len, off, nqt + 1, n_launches);
        const uint64_t H = read_u64(off + nqt, st);
        o.n_hits = H;
        o.hit_qid = keep.alloc<uint32_t>(H);
        o.hit_pid = keep.alloc<uint32_t>(H);
        o.hit_qpos = keep.alloc<uint32_t>(H);
        o.hit_tpos = keep.alloc<uint32_t>(H);
        o.hit_hash = keep.alloc<uint64_t>(H);
        if (H) {
            expand_hits_kernel<<<blocks_for(H), TB, 0, st>>>(csr, q_hash, q_loc, nqt, row0, off, H, o.hit_qid, o.hit_pid,
                                                            o.hit_hash, o.hit_qpos, o.hit_tpos);
            KS_CUDA(cudaGetLastError());
        }
        if (n_launches) *n_launches += 2;
    }
    *out = o;
}

}  // namespace ks
