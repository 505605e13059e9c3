// Dense k-mer space index build (hp alphabet, small k, scaled == 1): internal interface.  See sketch_dense_kernel
// (sketch.cu) for the idea; this file holds the per-handle rank tables and the CSR construction from sorted keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

// Scratch the table build needs next to its outputs (hashes + codes of all 2^k patterns, twice, + library sort space).
size_t dense_table_temp_bytes(uint32_t k);
// Per-handle tables of the path, in two steps with one host read in between (the code width depends on the data):
//   dense_build_tables: sorted_hash[2^k] (the patterns' MurmurHash3 values in increasing order), pattern_of_rank[2^k],
//     group_base[2^16 + 1] (first entry of every 16-bit hash-prefix group), d_flags (u32[4]): [0] != 0 when two patterns share
//     a hash or one hashes to 0 (the path must not be used then), [3] = largest rank of a pattern inside its prefix group;
//   dense_build_codes: code_of_pattern[2^k] = prefix << rb | (2 x rank in the group + 1), or prefix << rb | rank in the group
//     when the keys have no room for the parity bit (DenseSketchArgs, sketch.cuh).  Enqueued on `stream`; no synchronisation.
cudaError_t dense_build_tables(uint32_t k, uint64_t* sorted_hash, uint32_t* group_base, uint32_t* pattern_of_rank, void* temp,
                               size_t temp_bytes, uint32_t* d_flags, cudaStream_t stream, uint64_t* n_launches);
cudaError_t dense_build_codes(uint32_t k, const uint64_t* sorted_hash, const uint32_t* group_base, const uint32_t* pattern_of_rank,
                              int rb, int parity, uint32_t* code_of_pattern, cudaStream_t stream, uint64_t* n_launches);

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
    size_t off_cursor1 = 0, off_cursor2 = 0, off_chunks = 0, off_bstart = 0, off_counts = 0, off_overflow = 0;
    uint32_t max_chunks = 0;  // grid of the second level = entries of the chunk map (dense_chunks_kernel)
};
DenseSortPlan dense_sort_plan(uint64_t n, int rank_bits, int key_bits, int k);

struct DenseCsrArgs {
    DenseSortPlan plan;
    void* work = nullptr;  // plan.bytes; the rank kernel has scattered the keys into its first-level regions (plan.custom)
    uint64_t *keys_a, *keys_b;  // library sort only: packed keys in (protein, position) order in `a`; `b` scratch
    uint64_t n;
    uint32_t n_prot;
    uint32_t k;
    int rank_bits, pid_bits, pos_bits;  // rank_bits = DENSE_PREFIX_BITS + rb: the width of a code (DenseSketchArgs)
    int rb, parity;
    const uint64_t* offsets;      // device, n_prot + 1 (protein boundaries of the batch)
    const uint8_t* residues;      // device residues of the batch (bytes, or 5-bit codes when packed): the hash of an
    int packed;                   // exception key (even code) is recomputed from them (hp translation)
    const uint32_t* exc_flag;     // device: != 0 when the rank kernel emitted exception keys (picks the bucket kernel)
    const uint32_t* skip_flag;    // device: != 0 when the build is void (an exception the path does not handle)
    uint32_t* overflow = nullptr; // device: set when a sort bucket overflows (nullptr: the word in the plan's work buffer)
    int counts_zeroed = 0;        // the caller has zeroed d_counts on the stream already
    const uint64_t* sorted_hash;  // tables
    const uint32_t* group_base;
    // outputs
    uint64_t* loc;  // [n] postings (protein << 32 | position), ordered by (hash, protein, position)
    uint64_t* keys;
    uint32_t *key_grp, *grp_start, *t_size, *t_abund;
    uint64_t* d_counts;
    uint32_t* dir;  // hand-written sort: the bucket kernel fills it (segmented layout, [2^dir_bits + buckets + 1]); dir_bits >=
    int dir_bits;   // plan.total.  Library sort: compact layout, the caller runs launch_directory afterwards.
    // out (host): the layout written (index_build.cuh) and, when segmented, the bucket tables (device, inside `work`)
    int* out_dir_sub = nullptr;
    uint32_t* out_seg_nb = nullptr;
    const uint32_t** out_seg_start = nullptr;
    const uint64_t** out_seg_counts = nullptr;
    void* temp;  // dense_csr_temp_bytes(n)
    size_t temp_bytes;
    cudaEvent_t ev_sorted;  // recorded between the key sort and the CSR passes; may be null
};
size_t dense_csr_temp_bytes(uint64_t n);
// Library sort of the keys on their rank bits (stable), then two streaming passes: heads per tile, scan, CSR write.
cudaError_t dense_build_csr(const DenseCsrArgs& a, cudaStream_t stream, uint64_t* sort_launches, uint64_t* csr_launches);

}  // namespace ks
