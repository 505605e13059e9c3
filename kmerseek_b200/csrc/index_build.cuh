// Index build stage: sort of the (hash, loc) tuples + CSR construction.  Internal interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ks {

constexpr uint64_t MAX_TUPLES = (1ull << 31) - 1;  // per shard: two 31-bit counters share one scan word

// Device-resident CSR index over the sorted tuples.
//   hash[n], loc[n]            tuples ordered by (hash, protein, pos); loc = (protein << 32) | pos
//   keys[U]                    sorted unique hashes (= combined sketch mins)
//   key_grp[U+1]               first group of each key
//   grp_start[G+1]             first tuple of each (hash, protein) group; group size = abundance of the
//                              hash in that protein's sketch
//   t_size[P], t_abund[P]      per protein: distinct hashes (|T|) and kept windows (sum of abundances)
//   dir[2^dir_bits + 1]        bucket directory over the top hash bits: first key of each bucket
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
};

size_t sort_temp_bytes(uint64_t n, int end_bit);
// Sorts (hash, loc) pairs by hash, stable.  Input in (hash_a, loc_a); result is left in whichever
// pair *out_in_a says (1 = a, 0 = b).  LSD radix over bits [0, end_bit).
cudaError_t launch_sort(uint64_t* hash_a, uint64_t* loc_a, uint64_t* hash_b, uint64_t* loc_b, uint64_t n, int end_bit,
                        void* temp, size_t temp_bytes, cudaStream_t stream, int* out_in_a, uint64_t* n_launches);

// t_abund[p] = number of tuples of protein pid_base + p in a list ordered by (protein, pos).
cudaError_t launch_protein_abund(const uint64_t* loc, uint64_t n, uint32_t n_prot, uint32_t* t_abund, cudaStream_t stream,
                                 uint64_t* n_launches);

size_t csr_workspace_bytes(uint64_t n);
// Builds keys / key_grp / grp_start / t_size (t_size must hold t_abund on entry: non-head tuples are
// subtracted) and d_counts from sorted tuples, then the bucket directory.
cudaError_t launch_csr(const uint64_t* hash, const uint64_t* loc, uint64_t n, uint64_t* keys, uint32_t* key_grp,
                       uint32_t* grp_start, uint32_t* t_size, uint64_t* d_counts, uint32_t* dir, int dir_bits,
                       int dir_shift, void* workspace, cudaStream_t stream, uint64_t* n_launches);

}  // namespace ks
