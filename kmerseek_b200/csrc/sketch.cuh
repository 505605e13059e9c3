// Fused sketch stage: internal interface between the C ABI (api.cu) and the kernels (sketch.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dense_scatter.cuh"

namespace ks {

constexpr int SK_THREADS = 256;
constexpr int SK_ROWS = 8;                         // windows per thread
constexpr int SK_TILE = SK_THREADS * SK_ROWS;      // window start positions per CTA
constexpr int SK_MAX_TEMPLATE_K = 32;              // k above this takes the generic (byte-loop) kernel
constexpr int SK_MAX_K = 256;                      // bound of the generic kernel's halo

struct Lut256 { uint8_t b[256]; };

struct SketchArgs {
    const uint8_t* residues;   // device, 16 B aligned.  packed == 0: n_res bytes + >= 64 readable pad bytes;
                               // packed == 1: 5-bit codes, 8 residues per 5 bytes (pack_residues), + >= 72 pad bytes
    int packed = 0;
    const uint64_t* offsets;   // device, n_prot + 1
    uint64_t n_res;
    uint64_t n_prot;
    uint32_t k;
    int moltype;               // 0 protein (identity), 1 dayhoff, 2 hp
    uint64_t max_hash;
    uint32_t pid_base;         // added to the local protein index in the emitted tuples
    uint64_t* out_hash;        // device, capacity entries (already offset to the append position)
    uint64_t* out_loc;         // (pid << 32) | pos
    uint64_t capacity;
    uint64_t* d_count;         // device u64[2]: [0] tuples kept by this launch (may exceed capacity: nothing is written
                               // past it); [1] low word = ticket, high word != 0: a zero hash was met on the exact path
    PairScatter scatter{};     // scatter.out_key != nullptr: the tuples leave the kernel partitioned by the top bits of the
                               // hash (first level of an unstable partition, dense_scatter.cuh) instead of in (protein,
                               // pos) order; out_hash / out_loc / capacity are not used, d_count[0] still gets the total
    uint32_t* t_abund = nullptr;  // scatter mode with scaled > 1: kept windows per protein are counted here (device, n_prot,
                                  // zeroed by the caller): there is no (protein, pos)-ordered tuple array to count them from
    int force_general;         // 1: take the look-back path even when scaled == 1
    uint32_t tile_begin = 0, tile_end = 0;  // exact path: launch_sketch_tiles covers [tile_begin, tile_end); 0,0 = all
    void* workspace;           // sketch_workspace_bytes(n_res)
};

// Dense k-mer space path (hp, DENSE_MIN_K <= k <= DENSE_MAX_K, scaled == 1): see sketch_dense_kernel.
constexpr int DENSE_MIN_K = 8;
constexpr int DENSE_MAX_K = 24;
struct DenseSketchArgs {
    const uint32_t* rank_of_code;  // device, 2^k entries: rank of the pattern's hash among all patterns' hashes
    const uint64_t* sorted_hash;   // device, 2^k entries: the patterns' hashes in increasing order
    uint64_t* out_keys;            // device, capacity entries: rank' << (pid_bits + pos_bits) | protein << pos_bits | position
                                   // with rank' = 2 rank + 1 for a pattern; a window with a residue of neither class (no
                                   // pattern) is hashed from its bytes and gets rank' = 2 x (pattern hashes below its hash),
                                   // which sorts it between the right two patterns
    int pid_bits, pos_bits;
    int handle_exceptions;         // 0: such a window only raises exception_flag[0] (the caller takes the general path)
    uint32_t* exception_flag;      // device u32[2]: [0] unhandled exception / zero hash, [1] exception keys were emitted
    DenseScatter scatter;          // scatter.out != nullptr: the keys leave the kernel partitioned by their top bits
                                   // (first level of the key sort, dense_scatter.cuh) instead of in tile order
};
// launch_sketch_prepare first (exact path: scaled == 1); then this instead of launch_sketch_tiles.
cudaError_t launch_sketch_dense(const SketchArgs& a, const DenseSketchArgs& d, cudaStream_t stream, uint64_t* n_launches);

size_t sketch_workspace_bytes(uint64_t n_res);
// Enqueues memset + tile->protein map + the fused kernel on `stream`; adds the number of kernels launched.
cudaError_t launch_sketch(const SketchArgs& a, cudaStream_t stream, uint64_t* n_launches);
// The same in three steps, so that the host can stream the residues in while earlier tiles are hashed:
// prepare needs only the offsets; tiles [tile_begin, tile_end) need the residues up to tile_end * SK_TILE + k.
cudaError_t launch_sketch_prepare(const SketchArgs& a, cudaStream_t stream, uint64_t* n_launches);
cudaError_t launch_sketch_tiles(const SketchArgs& a, cudaStream_t stream, uint64_t* n_launches);
cudaError_t launch_sketch_finish(const SketchArgs& a, cudaStream_t stream);
bool sketch_is_exact(const SketchArgs& a);
void fill_lut(int moltype, Lut256* lut);
// 5-bit residue codes for the packed upload format: 'A'..'Z' -> 1..26, '*' -> 27, 0 = padding.  Returns false (and
// leaves `out` unspecified) when a byte has no code.  out must hold packed_bytes(n) bytes.
bool pack_residues(const uint8_t* res, uint64_t n, uint8_t* out);
inline uint64_t packed_bytes(uint64_t n) { return (n + 7) / 8 * 5 + 72; }

}  // namespace ks
