// Dense k-mer space index build (hp alphabet, small k, scaled == 1): internal interface.  See sketch_dense_kernel
// (sketch.cu) for the idea; this file holds the per-handle rank tables and the CSR construction from sorted keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

// Scratch the table build needs next to its outputs (hashes + codes of all 2^k patterns, twice, + library sort space).
size_t dense_table_temp_bytes(uint32_t k);
// rank_of_code[2^k] (u32) and sorted_hash[2^k] (u64): the patterns' MurmurHash3 values in increasing order and every
// pattern's position in that order.  *d_bad (device u32, zeroed here) becomes non-zero when two patterns share a hash
// or one hashes to 0 (the path must not be used then).  Enqueued on `stream`; no synchronisation.
cudaError_t dense_build_tables(uint32_t k, uint32_t* rank_of_code, uint64_t* sorted_hash, void* temp, size_t temp_bytes,
                               uint32_t* d_bad, cudaStream_t stream, uint64_t* n_launches);

// Plan of the hand-written key sort (dense_scatter.cuh, dense_bucket_kernel): the top `total` = l1 + l2 key bits pick
// one of 2^total final buckets of at most 4096 keys.  custom == 0: the input does not fit the scheme (more than
// 2^16 x 3072 keys, fewer rank bits than bucket bits, or more than 52 key bits below the bucket bits) and the library
// sorts the keys.
struct DenseSortPlan {
    int custom = 0;
    int l1 = 0, l2 = 0, total = 0;  // bits of the first level (fused into the rank kernel), the second, both
    uint32_t cap1 = 0;              // keys per first-level region
    // carve of the work buffer (bytes from its start)
    size_t off_region1 = 0, off_region2 = 0, off_small = 0, small_bytes = 0, bytes = 0;
    size_t off_cursor1 = 0, off_cursor2 = 0, off_chunks = 0, off_bstart = 0, off_status = 0, off_ticket = 0, off_overflow = 0;
};
DenseSortPlan dense_sort_plan(uint64_t n, int rank_bits, int key_bits);

struct DenseCsrArgs {
    DenseSortPlan plan;
    void* work = nullptr;  // plan.bytes; the rank kernel has scattered the keys into its first-level regions (plan.custom)
    uint64_t *keys_a, *keys_b;  // library sort only: packed keys in (protein, position) order in `a`; `b` scratch
    uint64_t n;
    uint32_t n_prot;
    uint32_t k;
    int rank_bits, pid_bits, pos_bits;  // rank_bits = k + 1: rank' = 2 rank + 1 (pattern) or 2 x lower bound (exception)
    const uint64_t* offsets;      // device, n_prot + 1 (protein boundaries of the batch)
    const uint8_t* residues;      // device residues of the batch (bytes, or 5-bit codes when packed): the hash of an
    int packed;                   // exception key (even rank') is recomputed from them (hp translation)
    const uint32_t* exc_flag;     // device: != 0 when the rank kernel emitted exception keys (picks the bucket kernel)
    const uint32_t* skip_flag;    // device: != 0 when the build is void (an exception the path does not handle)
    const uint64_t* sorted_hash;  // table
    // outputs
    uint64_t* loc;  // [n] postings (protein << 32 | position), ordered by (hash, protein, position)
    uint64_t* keys;
    uint32_t *key_grp, *grp_start, *t_size, *t_abund;
    uint64_t* d_counts;
    void* temp;  // dense_csr_temp_bytes(n)
    size_t temp_bytes;
    cudaEvent_t ev_sorted;  // recorded between the key sort and the CSR passes; may be null
};
size_t dense_csr_temp_bytes(uint64_t n);
// Library sort of the keys on their rank bits (stable), then two streaming passes: heads per tile, scan, CSR write.
cudaError_t dense_build_csr(const DenseCsrArgs& a, cudaStream_t stream, uint64_t* sort_launches, uint64_t* csr_launches);

}  // namespace ks
