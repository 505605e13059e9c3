// Dense k-mer space index build for sm_100a: rank tables and CSR construction from sorted rank keys.
//
// For the two-letter hp alphabet a k-mer is a k-bit pattern; with k <= 24 all 2^k patterns can be hashed once per
// handle.  Sorting those hashes gives every pattern its rank, and a (hash, protein, position) tuple collapses into one
// 64-bit key  rank | protein | position.  Keys leave the sketch kernel in (protein, position) order; a stable library
// sort on the rank bits alone puts them in (hash, protein, position) order -- equal ranks are equal hashes, so there is
// no bucket sort -- and two streaming passes turn the sorted keys into the same CSR arrays the general path produces
// (replaces, like index_build.cu, src/rust/index.rs:770-780 and :824-827 of the reference on the hot path).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "dense.cuh"

namespace ks {

namespace {

// MurmurHash3_x64_128 (seed 42, low word) of the k bytes 'h' / 'p' spelled by the low k bits of `pattern` (bit j set =
// byte j is 'h'); k <= 24.  Same arithmetic as the sketch kernels, on whole 64-bit words.
__device__ uint64_t murmur_of_pattern(uint32_t pattern, uint32_t k) {
    uint64_t w[3] = {0, 0, 0};
    for (uint32_t j = 0; j < k; j++) w[j >> 3] |= (uint64_t)(((pattern >> j) & 1u) ? 'h' : 'p') << (8 * (j & 7));
    uint64_t h1 = SEED, h2 = SEED;
    const uint32_t nb = k / 16, rem = k % 16;
    if (nb) {
        h1 ^= mix_k1(w[0]);
        h1 = rotl64(h1, 27) + h2;
        h1 = h1 * 5 + 0x52dce729;
        h2 ^= mix_k2(w[1]);
        h2 = rotl64(h2, 31) + h1;
        h2 = h2 * 5 + 0x38495ab5;
    }
    const uint64_t t1 = w[2 * nb], t2 = nb ? 0 : w[1];  // tail words (k <= 24: a second tail word only without a block)
    if (rem > 8) h2 ^= mix_k2(t2);
    if (rem > 0) h1 ^= mix_k1(t1);
    h1 ^= k;
    h2 ^= k;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    return h1 + h2;
}

__global__ void dense_hash_codes_kernel(uint32_t k, uint64_t* __restrict__ hash, uint32_t* __restrict__ code) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (1u << k)) return;
    hash[c] = murmur_of_pattern(c, k);
    code[c] = c;
}

__global__ void dense_rank_kernel(uint32_t n, const uint64_t* __restrict__ sorted_hash, const uint32_t* __restrict__ code_of_rank,
                                  uint32_t* __restrict__ rank_of_code, uint32_t* __restrict__ bad) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t h = sorted_hash[r];
    if (h == 0 || (r && sorted_hash[r - 1] == h)) atomicOr(bad, 1u);
    rank_of_code[code_of_rank[r]] = r;
}

// every complete window is a tuple (scaled == 1): kept windows per protein; distinct hashes start from the same number
// and lose one for every repeated (hash, protein) pair met while the groups are written
__global__ void dense_windows_kernel(const uint64_t* __restrict__ offsets, uint32_t n_prot, uint32_t k,
                                     uint32_t* __restrict__ t_abund, uint32_t* __restrict__ t_size) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_prot) return;
    const uint64_t len = offsets[p + 1] - offsets[p];
    const uint32_t w = len >= k ? (uint32_t)(len - k + 1) : 0u;
    t_abund[p] = w;
    t_size[p] = w;
}

constexpr int DC_THREADS = 256;
constexpr int DC_ROWS = 8;  // rows of 32 consecutive keys per warp
constexpr int DC_TILE = DC_THREADS * DC_ROWS;

// heads of a tile of sorted keys: a key head where the rank changes, a group head where (rank, protein) changes
__global__ void __launch_bounds__(DC_THREADS)
dense_count_kernel(const uint64_t* __restrict__ keys, uint64_t n, int loc_bits, int pos_bits, uint64_t* __restrict__ tile_counts) {
    __shared__ uint32_t s_k[DC_THREADS / 32], s_g[DC_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * DC_TILE;
    uint32_t ck = 0, cg = 0;
#pragma unroll
    for (int r = 0; r < DC_ROWS; r++) {
        const uint64_t i = base + (uint64_t)(warp * DC_ROWS + r) * 32 + lane;
        bool hk = false, hg = false;
        const uint64_t key = i < n ? keys[i] : 0ull;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (i < n) {
            if (lane == 0) prev = i ? keys[i - 1] : ~key;
            hk = (key >> loc_bits) != (prev >> loc_bits);
            hg = (key >> pos_bits) != (prev >> pos_bits);
        }
        ck += __popc(__ballot_sync(0xffffffffu, hk));
        cg += __popc(__ballot_sync(0xffffffffu, hg));
    }
    if (lane == 0) { s_k[warp] = ck; s_g[warp] = cg; }
    __syncthreads();
    if (tid == 0) {
        uint32_t k = 0, g = 0;
        for (int w = 0; w < DC_THREADS / 32; w++) { k += s_k[w]; g += s_g[w]; }
        tile_counts[blockIdx.x] = (uint64_t)k | ((uint64_t)g << 32);
        if (blockIdx.x == gridDim.x - 1) tile_counts[gridDim.x] = 0;
    }
}

__global__ void __launch_bounds__(DC_THREADS)
dense_write_kernel(const uint64_t* __restrict__ keys_in, uint64_t n, int loc_bits, int pos_bits, const uint64_t* __restrict__ tile_prefix,
                   const uint64_t* __restrict__ sorted_hash, uint64_t* __restrict__ loc, uint64_t* __restrict__ keys,
                   uint32_t* __restrict__ key_grp, uint32_t* __restrict__ grp_start, uint32_t* __restrict__ t_size,
                   uint64_t* __restrict__ d_counts) {
    __shared__ uint32_t s_k[DC_THREADS / 32], s_g[DC_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t base = (uint64_t)blockIdx.x * DC_TILE;
    const uint64_t pos_mask = (1ull << pos_bits) - 1ull, pid_mask = (1ull << (loc_bits - pos_bits)) - 1ull;
    uint64_t key[DC_ROWS];
    uint32_t bk[DC_ROWS], bg[DC_ROWS];
    uint32_t wk = 0, wg = 0;
#pragma unroll
    for (int r = 0; r < DC_ROWS; r++) {
        const uint64_t i = base + (uint64_t)(warp * DC_ROWS + r) * 32 + lane;
        bool hk = false, hg = false;
        key[r] = i < n ? keys_in[i] : 0ull;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key[r], 1);
        if (i < n) {
            if (lane == 0) prev = i ? keys_in[i - 1] : ~key[r];
            hk = (key[r] >> loc_bits) != (prev >> loc_bits);
            hg = (key[r] >> pos_bits) != (prev >> pos_bits);
            const uint32_t pid = (uint32_t)((key[r] >> pos_bits) & pid_mask);
            loc[i] = ((uint64_t)pid << 32) | (key[r] & pos_mask);
            if (!hg) atomicSub(&t_size[pid], 1u);  // the hash again in the same protein: one distinct hash fewer
        }
        bk[r] = __ballot_sync(0xffffffffu, hk);
        bg[r] = __ballot_sync(0xffffffffu, hg);
        wk += __popc(bk[r]);
        wg += __popc(bg[r]);
    }
    if (lane == 0) { s_k[warp] = wk; s_g[warp] = wg; }
    __syncthreads();
    const uint64_t pfx = tile_prefix[blockIdx.x];
    uint32_t rk = (uint32_t)pfx, rg = (uint32_t)(pfx >> 32);
    uint32_t tk = 0, tg = 0;
#pragma unroll
    for (int w = 0; w < DC_THREADS / 32; w++) {
        if (w < (int)warp) { rk += s_k[w]; rg += s_g[w]; }
        tk += s_k[w];
        tg += s_g[w];
    }
#pragma unroll
    for (int r = 0; r < DC_ROWS; r++) {
        const uint64_t i = base + (uint64_t)(warp * DC_ROWS + r) * 32 + lane;
        const uint32_t g = rg + __popc(bg[r] & lt);
        if ((bg[r] >> lane) & 1u) grp_start[g] = (uint32_t)i;
        if ((bk[r] >> lane) & 1u) {
            const uint32_t u = rk + __popc(bk[r] & lt);
            keys[u] = sorted_hash[key[r] >> loc_bits];
            key_grp[u] = g;
        }
        rk += __popc(bk[r]);
        rg += __popc(bg[r]);
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        const uint64_t U = (uint32_t)pfx + tk, G = (uint32_t)(pfx >> 32) + tg;
        d_counts[0] = U;
        d_counts[1] = G;
        key_grp[U] = (uint32_t)G;
        grp_start[G] = (uint32_t)n;
    }
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t table_sort_bytes(uint32_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int64_t)n, 0, 64);
    return align256(b);
}

size_t key_sort_bytes(uint64_t n) {
    size_t b = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, b, k, (int64_t)n, 0, 64);
    return align256(b);
}

size_t scan_bytes(uint64_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int64_t)n);
    return align256(b);
}

}  // namespace

size_t dense_table_temp_bytes(uint32_t k) {
    const size_t n = (size_t)1 << k;
    return align256(n * 8) + 2 * align256(n * 4) + table_sort_bytes((uint32_t)n) + 256;
}

cudaError_t dense_build_tables(uint32_t k, uint32_t* rank_of_code, uint64_t* sorted_hash, void* temp, size_t temp_bytes,
                               uint32_t* d_bad, cudaStream_t stream, uint64_t* n_launches) {
#define KS_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
    const uint32_t n = 1u << k;
    if (temp_bytes < dense_table_temp_bytes(k)) return cudaErrorInvalidValue;
    char* p = (char*)temp;
    uint64_t* hash = (uint64_t*)p; p += align256((size_t)n * 8);
    uint32_t* code = (uint32_t*)p; p += align256((size_t)n * 4);
    uint32_t* code_of_rank = (uint32_t*)p; p += align256((size_t)n * 4);
    size_t sort_bytes = table_sort_bytes(n);
    KS_TRY(cudaMemsetAsync(d_bad, 0, 4, stream));
    dense_hash_codes_kernel<<<(n + 255) / 256, 256, 0, stream>>>(k, hash, code);
    KS_TRY(cudaGetLastError());
    KS_TRY(cub::DeviceRadixSort::SortPairs(p, sort_bytes, hash, sorted_hash, code, code_of_rank, (int64_t)n, 0, 64, stream));
    dense_rank_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sorted_hash, code_of_rank, rank_of_code, d_bad);
    if (n_launches) *n_launches += 2 + 10;
    return cudaGetLastError();
}

size_t dense_csr_temp_bytes(uint64_t n) {
    const uint64_t nt = (n + DC_TILE - 1) / DC_TILE;
    return key_sort_bytes(n) + 2 * align256((nt + 1) * 8) + scan_bytes(nt + 1) + 256;
}

cudaError_t dense_build_csr(const DenseCsrArgs& a, cudaStream_t stream, uint64_t* sort_launches, uint64_t* csr_launches) {
    const uint64_t n = a.n;
    const int loc_bits = a.pid_bits + a.pos_bits;
    if (a.n_prot) {
        dense_windows_kernel<<<(a.n_prot + 255) / 256, 256, 0, stream>>>(a.offsets, a.n_prot, a.k, a.t_abund, a.t_size);
        KS_TRY(cudaGetLastError());
        if (csr_launches) *csr_launches += 1;
    }
    if (n == 0) {
        KS_TRY(cudaMemsetAsync(a.d_counts, 0, 16, stream));
        KS_TRY(cudaMemsetAsync(a.key_grp, 0, 4, stream));
        KS_TRY(cudaMemsetAsync(a.grp_start, 0, 4, stream));
        if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
        return cudaSuccess;
    }
    char* p = (char*)a.temp;
    size_t sort_bytes = key_sort_bytes(n);
    void* sort_temp = p; p += sort_bytes;
    const uint64_t nt = (n + DC_TILE - 1) / DC_TILE;
    uint64_t* tile_counts = (uint64_t*)p; p += align256((nt + 1) * 8);
    uint64_t* tile_prefix = (uint64_t*)p; p += align256((nt + 1) * 8);
    size_t sbytes = scan_bytes(nt + 1);
    void* scan_temp = p;
    // 1. stable library sort on the rank bits: keys arrive in (protein, position) order and leave in (rank, protein, position)
    cub::DoubleBuffer<uint64_t> kb(a.keys_a, a.keys_b);
    KS_TRY(cub::DeviceRadixSort::SortKeys(sort_temp, sort_bytes, kb, (int64_t)n, loc_bits, loc_bits + a.rank_bits, stream));
    if (sort_launches) *sort_launches += 2 + (a.rank_bits + 7) / 8;
    if (a.ev_sorted) KS_TRY(cudaEventRecord(a.ev_sorted, stream));
    const uint64_t* sorted = kb.Current();
    // 2. heads per tile, scan, CSR write
    dense_count_kernel<<<(unsigned)nt, DC_THREADS, 0, stream>>>(sorted, n, loc_bits, a.pos_bits, tile_counts);
    KS_TRY(cudaGetLastError());
    KS_TRY(cub::DeviceScan::ExclusiveSum(scan_temp, sbytes, tile_counts, tile_prefix, (int64_t)(nt + 1), stream));
    dense_write_kernel<<<(unsigned)nt, DC_THREADS, 0, stream>>>(sorted, n, loc_bits, a.pos_bits, tile_prefix, a.sorted_hash, a.loc,
                                                                a.keys, a.key_grp, a.grp_start, a.t_size, a.d_counts);
    if (csr_launches) *csr_launches += 4;
    return cudaGetLastError();
#undef KS_TRY
}

}  // namespace ks
