// Index build stage: sort of the (hash, loc) tuples + CSR construction.  Internal interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

constexpr uint64_t MAX_TUPLES = (1ull << 31) - 1;  // per shard: 32-bit postings offsets, 31-bit scan halves

// Device-resident CSR index over the sorted tuples.
//   hash[n], loc[n]            tuples ordered by (hash, protein, pos); loc = (protein << 32) | pos
//   keys[U]                    sorted unique hashes (= combined sketch mins)
//   key_grp[U+1]               first group of each key
//   grp_start[G+1]             first tuple of each (hash, protein) group; group size = abundance of the
//                              hash in that protein's sketch
//   t_size[P], t_abund[P]      per protein: distinct hashes (|T|) and kept windows (sum of abundances)
//   dir[2^dir_bits + 1]        bucket directory over the top hash bits: first key of each bucket
//
// Two layouts of keys / key_grp / grp_start (the postings loc[] are dense and ordered in both):
//   compact    keys[0 .. U), groups [0 .. G): what the streaming passes (scan_counts + csr_write) write.
//   segmented  (dir_sub < 31) what the bucket-sort kernels write themselves, WITHOUT any chain between buckets: sort
//              bucket b -- tuples [seg_start[b], seg_start[b + 1]) -- keeps its keys at positions seg_start[b] + b + i and
//              its groups at the same base (a bucket of m tuples has at most m keys and m groups, so the segments never
//              meet; the spare slot per bucket holds the sentinel), key_grp / grp_start hold positions in that space, and
//              every bucket has its own 2^dir_sub + 1 directory entries.  A reader only ever goes from a key position
//              to u + 1, from a group position to g + 1: both layouts read the same way once find_key has mapped a hash to
//              its directory slot (x + (x >> dir_sub)).  seg_counts[b] = keys | groups << 32 of bucket b (for exports).
struct CsrView {
    const uint64_t* hash;
    const uint64_t* loc;
    const uint64_t* keys;
    const uint32_t* key_grp;
    const uint32_t* grp_start;
    const uint32_t* t_size;
    const uint32_t* t_abund;
    const uint32_t* dir;
    const uint64_t* d_counts;  // [0] = U, [1] = G  (device)
    uint64_t n;
    uint32_t n_prot;
    int dir_bits;
    int dir_shift;
    int dir_sub = 31;                        // 31: compact layout; else log2 of the directory entries per sort bucket
    uint32_t seg_nb = 0;                     // segmented: sort buckets
    const uint32_t* seg_start = nullptr;     // [seg_nb + 1]
    const uint64_t* seg_counts = nullptr;    // [seg_nb]
};
constexpr int DIR_SUB_COMPACT = 31;

// Plan of the unstable partition of the general path (dense_scatter.cuh): the top `total` = l1 + l2 bits of the normalised
// hash pick one of 2^total buckets of at most 4096 tuples; the first level is fused into the sketch kernel.  custom == 0:
// the input does not fit (fewer than 2^16 tuples, more than 16 bucket bits) and the stable library partition is used.
struct PairSortPlan {
    int custom = 0;
    int l1 = 0, l2 = 0, total = 0;
    uint32_t cap1 = 0;  // tuples per first-level region
    // carve of the work buffer
    size_t off_r1_hash = 0, off_r1_loc = 0, off_r2_hash = 0, off_r2_loc = 0, off_small = 0, small_bytes = 0;
    size_t off_cursor1 = 0, off_cursor2 = 0, off_overflow = 0, off_chunks = 0, off_bstart = 0, bytes = 0;
    uint32_t max_chunks = 0;  // grid of the second level = entries of the chunk map (dense_chunks_kernel)
};
PairSortPlan pair_sort_plan(uint64_t n, int end_bit, uint64_t max_hash);

struct BuildArgs {
    // scattered input (plan.custom): the sketch kernel has put the tuples into the first-level regions of `work`; the
    // postings are written to loc_a.
    PairSortPlan plan;
    void* work = nullptr;
    const uint64_t* offsets = nullptr;  // device, n_prot + 1
    uint32_t k = 0;
    const uint32_t** overflow_dev = nullptr;  // out: device flag, != 0 after the build when a region overflowed (the build
                                              // must then be redone from ordered tuples); the caller reads it
    int abund_ready = 0;                // t_abund was filled by the sketch kernel (scaled > 1); else it comes from the offsets
    // tuples in (protein, pos) order in the `a` pair; `b` is scratch of the same size.
    uint64_t *hash_a, *loc_a, *hash_b, *loc_b;
    uint64_t n;
    uint32_t n_prot;
    int end_bit;  // hashes are < 2^end_bit (64 - leading zeros of max_hash)
    uint64_t max_hash;
    int repeat_heavy;  // the k-mer space is small next to n (hashes repeat many times): two local counting passes
    double avg_postings = 0.0;  // tuples per possible hash (n / (alphabet^k / scaled)): above ~1000 a single hash fills a
                                // bucket and the build takes the library sort of all bits
    int ls_variant = 0;  // test hook: 0 pick from repeat_heavy, 1 rep, 2 bin, 3 bin with a barrier per row
    // outputs (device, preallocated): keys[n], key_grp[n+1], grp_start[n+1], t_size[P], t_abund[P], d_counts[2],
    // dir[2^dir_bits + 1]
    uint64_t* keys;
    uint32_t *key_grp, *grp_start, *t_size, *t_abund;
    uint64_t* d_counts;
    uint32_t* dir;        // [2^dir_bits + max buckets + 1]
    int dir_bits, dir_shift;  // dir_bits >= build_top_bits(): a directory bucket never spans two sort buckets
    // out (host): layout of what was written -- DIR_SUB_COMPACT, or the segmented layout's dir_sub with its bucket tables
    // (device pointers into `temp` / `work`, valid until the next build; the caller keeps a copy)
    int* out_dir_sub = nullptr;
    uint32_t* out_seg_nb = nullptr;
    const uint32_t** out_seg_start = nullptr;
    const uint64_t** out_seg_counts = nullptr;
    void* temp;
    size_t temp_bytes;
    int* hash_written = nullptr;  // out (host): 0 when the sorted hash column was not written (see expand_sorted_hash)
    cudaEvent_t ev_partitioned;  // recorded after the partition by the top bits (stage timing); may be null
    cudaEvent_t ev_sorted;       // recorded between the sort and the CSR write; may be null
};

size_t build_temp_bytes(uint64_t n, int end_bit);
// Sort buckets the build of n ordered tuples uses (2^bits), or -1 when it takes the library sort for all bits (compact
// layout); and the slack the key / group arrays need past n (one sentinel slot per bucket).
int build_top_bits(uint64_t n, int end_bit, uint64_t max_hash, double avg_postings);
uint64_t build_slack(uint64_t n);
// Sort by hash (stable) + CSR build + directory.  *out_in_a = 1 when the sorted tuples ended in the `a` pair.
// Scattered input: no synchronisation (the caller reads *overflow_dev with the totals); ordered input: synchronises the
// stream once (oversize-bucket count).  Adds the kernels launched to the two counters.
cudaError_t build_index(const BuildArgs& a, cudaStream_t stream, int* out_in_a, uint64_t* sort_launches,
                        uint64_t* csr_launches);

// dir[x] = index of the first key whose top dir_bits bits (of the normalised hash) are >= x; dir[2^dir_bits] = U.
cudaError_t launch_directory(const uint64_t* keys, const uint64_t* d_counts, uint32_t* dir, int dir_bits, int dir_shift,
                             cudaStream_t stream);

// Rebuild hash[i] of the sorted tuples from keys / key_grp / grp_start (the fused bucket sort does not write the
// column: nothing on the hot path reads it).
cudaError_t expand_sorted_hash(const CsrView& v, uint64_t* hash, cudaStream_t stream);

}  // namespace ks
