// Query stage: internal interface (api.cu <-> search.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "index_build.cuh"
#include "sketch.cuh"
#include "util.cuh"

namespace ks {

// Tuples grouped into per-owner sketches (owner = protein or query): sorted by (owner, hash, pos).
struct Grouped {
    uint64_t n = 0;          // tuples
    uint64_t n_entries = 0;  // distinct (owner, hash)
    uint64_t* s_hash = nullptr;    // [n]   sorted tuples
    uint64_t* s_loc = nullptr;     // [n]
    uint64_t* ent_hash = nullptr;  // [E]   = concatenated per-owner mins
    uint32_t* ent_owner = nullptr; // [E]   local owner index
    uint32_t* ent_first = nullptr; // [E+1] first sorted tuple of the entry; diff = abundance
    uint64_t* sig_ptr = nullptr;   // [n_owner+1] CSR over entries
};

// `hash`/`loc` are left untouched.  Allocations come from `keep`; scratch from `tmp`.  (Export path: library sorts.)
void group_by_owner(Arena& keep, Arena& tmp, const uint64_t* hash, const uint64_t* loc, uint64_t n, uint32_t n_owner,
                    int hash_end_bit, Grouped* out, uint64_t* n_launches);

constexpr int N_SCORE_COLS = 12;
// order of the score columns
enum ScoreCol {
    SC_CONTAINMENT = 0, SC_CONTAINMENT_TARGET, SC_MAX_CONTAINMENT, SC_JACCARD, SC_QUERY_ANI, SC_MATCH_ANI,
    SC_AVERAGE_ANI, SC_MAX_ANI, SC_AVERAGE_ABUND, SC_MEDIAN_ABUND, SC_STD_ABUND, SC_F_WEIGHTED
};
constexpr int N_PAIR_U32 = 5;  // pair_qid, pair_pid, intersect_hashes, q_size, t_size
constexpr int N_PAIR_U64 = 2;  // n_weighted_found, total_weighted_hashes
constexpr int N_HIT_U32 = 4;   // hit_qid, hit_pid, hit_qpos, hit_tpos
constexpr uint64_t PAIR_ROW_BYTES = N_PAIR_U32 * 4 + N_PAIR_U64 * 8 + N_SCORE_COLS * 8;  // 132
constexpr uint64_t HIT_ROW_BYTES = N_HIT_U32 * 4 + 8;                                    // 24

// One search result = ONE contiguous block (device, and its pinned host copy with the same layout): the per-query CSR
// over the query sketches, the pair columns, then (optional) the hit columns and the query sketches.  Every column starts
// on a 256-byte boundary.  One cudaMemcpyAsync brings a result to the host; one ncclSend ships a shard's result.
struct ResultLayout {
    uint64_t nq = 0, n_pairs = 0, n_hits = 0, n_entries = 0;
    int hits = 0, sketches = 0;
    int wire = 0;                                 // sharded search: the merge offsets travel with the block
    uint64_t n_res = 0;                           // residues of the query batch (wire + hits: length of win_hoff)
    size_t off_sig_ptr = 0;                       // u64[nq + 1]
    size_t off_u32[N_PAIR_U32] = {};              // u32[n_pairs] each
    size_t off_u64[N_PAIR_U64] = {};              // u64[n_pairs] each
    size_t off_score[N_SCORE_COLS] = {};          // f64[n_pairs] each
    size_t off_hit32[N_HIT_U32] = {};             // u32[n_hits] each
    size_t off_hit_hash = 0;                      // u64[n_hits]
    size_t off_q_mins = 0, off_q_abunds = 0;      // u64[n_entries] each
    size_t off_pair_off = 0, off_hit_off = 0;     // wire: u64[nq + 1] exclusive pair / hit offsets per query
    size_t off_win_hoff = 0;                      // wire + hits: u32[n_res] hits before window w inside its query
    size_t bytes = 0;
};
ResultLayout result_layout(uint64_t nq, uint64_t n_pairs, uint64_t n_hits, uint64_t n_entries, bool hits, bool sketches,
                           bool wire = false, uint64_t n_res = 0);

// A scored pair as the query kernel leaves it (compact staging; the finalize kernel expands it into the 19 columns).
struct StagedPair {
    uint32_t qid, rank;   // rank of the pair among its query's pairs (targets ascending)
    uint32_t pid, isect;  // shard-local protein id, |Q n T|
    uint64_t sum_a;       // sum of the target's abundances over the intersection
    double median, stdev; // of those abundances (population)
};

// Limits of the hand-written path: a query of more windows takes the library-sorted path (search_device_legacy).
constexpr uint32_t QK_SMALL_WINDOWS = 512;
constexpr uint32_t QK_MAX_WINDOWS = 4096;

// Device scratch of the query path (owned by the handle, grow-only).  Sparse arrays are indexed by the residue offset of
// the query (offs[q] + i): a query has at most as many windows / distinct hashes as residues.
struct QueryScratch {
    uint32_t *e_count = nullptr, *p_count = nullptr;  // [nq] distinct hashes |Q|, scored pairs
    uint64_t* h_count = nullptr;                      // [nq] hits
    uint64_t *sig_ptr = nullptr, *pair_off = nullptr, *hit_off = nullptr;  // [nq + 1] exclusive scans of the three
    uint64_t* ent_hash = nullptr;                     // [n_res] sorted distinct hashes of query q at offs[q] ..
    uint32_t* ent_abund = nullptr;                    // [n_res] their abundances
    uint32_t *win_key = nullptr, *win_hoff = nullptr; // [n_res] per window: index key (or ~0), hits before it in its query
    StagedPair* stage = nullptr;
    uint64_t stage_cap = 0;
    uint64_t* totals = nullptr;  // device u64[8]: [0] entries, [1] pairs, [2] hits, [3] staging cursor, [4] error flags
};
enum { QT_ENTRIES = 0, QT_PAIRS = 1, QT_HITS = 2, QT_CURSOR = 3, QT_FLAGS = 4, QT_WORDS = 8 };
constexpr uint64_t QF_HITS_OVERFLOW = 1;  // one query has 2^32 hits or more

struct QueryBatchView {
    const uint8_t* res;    // device, plain bytes (+ >= 64 readable pad bytes)
    const uint64_t* offs;  // device, nq + 1
    uint32_t nq;
    uint64_t n_res;
    uint32_t max_windows;  // longest query, in windows (host-known from the offsets)
};

// Phase 1 (no host round trip inside): per-query CTAs sketch the query (translate, hash, filter, sort, distinct),
// look every distinct hash up in the index, expand the (target, abundance) records behind the found keys, sort them and
// leave one StagedPair per (query, target); then one scan over the queries.  totals[] is complete when the stream
// reaches the end of this call; pairs past stage_cap are counted but not stored (the caller grows the staging buffer and
// calls again).
cudaError_t launch_query_phase1(const QueryBatchView& q, const CsrView& csr, uint32_t k, int moltype, uint64_t max_hash,
                                bool want_hits, const QueryScratch& s, cudaStream_t stream, uint64_t* n_launches);
// Phase 2: staged pairs -> the 19 pair columns at (query, target) order; hit list; compact query sketches; sig_ptr.
// `block` is a device block of layout `L` (n_pairs / n_hits / n_entries as read back from totals[]).
cudaError_t launch_query_phase2(const QueryBatchView& q, const CsrView& csr, uint32_t k, uint32_t pid_base,
                                const QueryScratch& s, const ResultLayout& L, void* block, cudaStream_t stream,
                                uint64_t* n_launches);

// Library-sorted path for batches with a query of more than QK_MAX_WINDOWS windows: the round-1 pipeline (sketch kernel of
// the build + CUB sorts and scans), written into the same block layout.  q_hash/q_loc: query tuples in (query, qpos)
// order.  Synchronises the stream several times.
struct LegacyCounts { uint64_t n_pairs = 0, n_hits = 0, n_entries = 0; };
void search_device_legacy(Arena& tmp, const CsrView& csr, const uint64_t* q_hash, const uint64_t* q_loc, uint64_t n_q_tuples,
                          uint32_t n_queries, uint32_t ksize, int hash_end_bit, bool want_hits, bool want_sketches,
                          uint32_t pid_base, Arena& block_owner, void** block_out, ResultLayout* layout_out,
                          uint64_t* n_launches);

// ---- multi-GPU merge (rank 0): shard results -> one result ordered by (query, target) -----------------------------
// Shards hold ascending protein ranges and every shard's pairs are ordered by (query, target), its hits by (query, qpos,
// target, tpos).  So row j of shard s lands at a position that is a sum of the shards' own exclusive offsets: with
// off_s[q] = first pair of query q in shard s,
//     dst = sum_s' off_s'[q]  +  sum_{s' < s} (off_s'[q + 1] - off_s'[q])  +  (j - off_s[q])
// and the same per (query, window) for hits.  Counting, not sorting; no scan on rank 0.
struct MergeShard {
    const void* block = nullptr;  // the shard's result block in wire layout (device memory of this rank)
    ResultLayout layout;
};
constexpr int MAX_SHARDS = 16;
struct MergeArgs {
    int n_shards = 0;
    MergeShard shard[MAX_SHARDS];
    uint32_t nq = 0;
    const uint64_t* q_offs = nullptr;  // device, nq + 1: residue offsets of the query batch
    uint32_t k = 0;
    void* out_block = nullptr;         // merged block; its sig_ptr / query sketches are copied from shard 0's block
    ResultLayout out_layout;
};
cudaError_t launch_merge(const MergeArgs& m, cudaStream_t stream, uint64_t* n_launches);

// sums[p] += every distinct hash of protein p (wrapping): the kmerseek id of a signature is the hex of that sum
// (src/rust/signature.rs:277-279).  `sums` (n_prot words) must be zeroed.  signature_count() support; not a hot path.
cudaError_t launch_id_sums(const CsrView& v, uint64_t n_keys, unsigned long long* sums, cudaStream_t stream);

}  // namespace ks
