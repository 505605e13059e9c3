// Query stage: internal interface (api.cu <-> search.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "index_build.cuh"
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

// `hash`/`loc` are left untouched.  Allocations come from `keep`; scratch from `tmp`.
void group_by_owner(Arena& keep, Arena& tmp, const uint64_t* hash, const uint64_t* loc, uint64_t n, uint32_t n_owner,
                    int hash_end_bit, Grouped* out, uint64_t* n_launches);

constexpr int N_SCORE_COLS = 12;
// order of SearchDevice::score[]
enum ScoreCol {
    SC_CONTAINMENT = 0, SC_CONTAINMENT_TARGET, SC_MAX_CONTAINMENT, SC_JACCARD, SC_QUERY_ANI, SC_MATCH_ANI,
    SC_AVERAGE_ANI, SC_MAX_ANI, SC_AVERAGE_ABUND, SC_MEDIAN_ABUND, SC_STD_ABUND, SC_F_WEIGHTED
};

struct SearchDevice {
    uint64_t n_pairs = 0;
    uint32_t *pair_qid = nullptr, *pair_pid = nullptr, *intersect = nullptr, *q_size = nullptr, *t_size = nullptr;
    uint64_t *n_weighted_found = nullptr, *total_weighted = nullptr;
    double* score[N_SCORE_COLS] = {};
    uint64_t n_hits = 0;
    uint32_t *hit_qid = nullptr, *hit_pid = nullptr, *hit_qpos = nullptr, *hit_tpos = nullptr;
    uint64_t* hit_hash = nullptr;
};

// q_hash/q_loc: query tuples in (query, qpos) order as the sketch kernel emits them.
void search_device(Arena& keep, Arena& tmp, const CsrView& csr, const uint64_t* q_hash, const uint64_t* q_loc,
                   uint64_t n_q_tuples, uint32_t n_queries, uint32_t ksize, int hash_end_bit, bool want_hits,
                   Grouped* q_sketches, SearchDevice* out, uint64_t* n_launches);

}  // namespace ks
