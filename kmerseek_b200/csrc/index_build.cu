// Index build for sm_100a: sort of the sketch tuples by hash and single-pass CSR construction.
//
// Replaces, on the hot path, the reference's per-protein sorted Vec inserts (sourmash KmerMinHash via
// src/rust/signature.rs:273-274), the nested HashMap position records (src/rust/index.rs:770-780) and the
// serial sorted-Vec merge into combined_minhash (src/rust/index.rs:824-827).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "index_build.cuh"

namespace ks {

namespace {

constexpr int CSR_THREADS = 256;
constexpr int CSR_ROWS = 8;
constexpr int CSR_TILE = CSR_THREADS * CSR_ROWS;

__global__ void protein_abund_kernel(const uint64_t* __restrict__ loc, uint64_t n, uint32_t n_prot,
                                     uint32_t* __restrict__ t_abund) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_prot) return;
    // tuples are ordered by (protein, pos): count = lower_bound((p+1)<<32) - lower_bound(p<<32)
    auto lb = [&](uint64_t key) {
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (loc[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    uint64_t a = lb((uint64_t)p << 32);
    uint64_t b = (p + 1 == 0) ? n : lb(((uint64_t)p + 1) << 32);
    t_abund[p] = (uint32_t)(b - a);
}

// One pass over the sorted tuples.  Two running counts (unique hashes, (hash, protein) groups) are packed
// into one scan word and chained across CTAs with the same decoupled look-back the sketch kernel uses.
__global__ void __launch_bounds__(CSR_THREADS)
csr_kernel(const uint64_t* __restrict__ hash, const uint64_t* __restrict__ loc, uint64_t n, uint64_t* __restrict__ keys,
           uint32_t* __restrict__ key_grp, uint32_t* __restrict__ grp_start, uint32_t* __restrict__ t_size,
           uint64_t* __restrict__ d_counts, uint32_t* __restrict__ ticket, uint64_t* __restrict__ status) {
    __shared__ uint32_t s_wk[CSR_THREADS / 32], s_wg[CSR_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t n_tiles = (n + CSR_TILE - 1) / CSR_TILE;
    const uint64_t i0 = (uint64_t)tile * CSR_TILE + warp * (CSR_ROWS * 32) + lane;

    uint64_t h[CSR_ROWS];
    uint32_t bk[CSR_ROWS], bg[CSR_ROWS];
    uint32_t wk = 0, wg = 0;
#pragma unroll
    for (int r = 0; r < CSR_ROWS; r++) {
        const uint64_t i = i0 + r * 32;
        bool hk = false, hg = false;
        h[r] = 0;
        if (i < n) {
            h[r] = hash[i];
            const uint32_t pid = (uint32_t)(loc[i] >> 32);
            if (i == 0) {
                hk = hg = true;
            } else {
                hk = h[r] != hash[i - 1];
                hg = hk || pid != (uint32_t)(loc[i - 1] >> 32);
            }
            if (!hg) atomicSub(&t_size[pid], 1u);  // a repeat of (hash, protein): not a new min of that sketch
        }
        bk[r] = __ballot_sync(0xffffffffu, hk);
        bg[r] = __ballot_sync(0xffffffffu, hg);
        wk += __popc(bk[r]);
        wg += __popc(bg[r]);
    }
    if (lane == 0) { s_wk[warp] = wk; s_wg[warp] = wg; }
    __syncthreads();
    uint32_t pk = 0, pg = 0, tk = 0, tg = 0;
#pragma unroll
    for (int w = 0; w < CSR_THREADS / 32; w++) {
        if (w < (int)warp) { pk += s_wk[w]; pg += s_wg[w]; }
        tk += s_wk[w];
        tg += s_wg[w];
    }
    if (warp == 0) {
        uint64_t excl = scan_lookback(status, tile, (uint64_t)tk | ((uint64_t)tg << 31));
        if (lane == 0) {
            s_base = excl;
            if (tile == n_tiles - 1) {
                uint64_t U = (excl & 0x7fffffffu) + tk, G = (excl >> 31) + tg;
                d_counts[0] = U;
                d_counts[1] = G;
                key_grp[U] = (uint32_t)G;
                grp_start[G] = (uint32_t)n;
            }
        }
    }
    __syncthreads();
    uint32_t rk = (uint32_t)(s_base & 0x7fffffffu) + pk;
    uint32_t rg = (uint32_t)(s_base >> 31) + pg;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < CSR_ROWS; r++) {
        const uint64_t i = i0 + r * 32;
        const bool hk = (bk[r] >> lane) & 1u, hg = (bg[r] >> lane) & 1u;
        const uint32_t g = rg + __popc(bg[r] & lt);
        if (hg) grp_start[g] = (uint32_t)i;
        if (hk) {
            const uint32_t u = rk + __popc(bk[r] & lt);
            keys[u] = h[r];
            key_grp[u] = g;
        }
        rk += __popc(bk[r]);
        rg += __popc(bg[r]);
    }
}

// dir[b] = index of the first key whose bucket is >= b; dir[2^bits] = U.
__global__ void dir_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ d_counts,
                           uint32_t* __restrict__ dir, int bits, int shift) {
    const uint64_t U = d_counts[0];
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint32_t nb = 1u << bits;
    if (U == 0) {
        for (uint64_t b = i; b <= nb; b += (uint64_t)gridDim.x * blockDim.x) dir[b] = 0;
        return;
    }
    if (i >= U) return;
    const uint32_t b = (uint32_t)(keys[i] >> shift);
    const int64_t bp = i ? (int64_t)(uint32_t)(keys[i - 1] >> shift) : -1;
    for (int64_t x = bp + 1; x <= (int64_t)b; x++) dir[x] = (uint32_t)i;
    if (i == U - 1)
        for (uint32_t x = b + 1; x <= nb; x++) dir[x] = (uint32_t)U;
}

}  // namespace

size_t sort_temp_bytes(uint64_t n, int end_bit) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int64_t)n, 0, end_bit);
    return bytes;
}

cudaError_t launch_sort(uint64_t* hash_a, uint64_t* loc_a, uint64_t* hash_b, uint64_t* loc_b, uint64_t n, int end_bit,
                        void* temp, size_t temp_bytes, cudaStream_t stream, int* out_in_a, uint64_t* n_launches) {
    cub::DoubleBuffer<uint64_t> k(hash_a, hash_b), v(loc_a, loc_b);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, (int64_t)n, 0, end_bit, stream);
    *out_in_a = k.Current() == hash_a ? 1 : 0;
    if (n_launches) *n_launches += 2 + (end_bit + 7) / 8;  // histogram + scan + one onesweep pass per 8 bits
    return e;
}

cudaError_t launch_protein_abund(const uint64_t* loc, uint64_t n, uint32_t n_prot, uint32_t* t_abund, cudaStream_t stream,
                                 uint64_t* n_launches) {
    if (n_prot == 0) return cudaSuccess;
    protein_abund_kernel<<<(n_prot + 255) / 256, 256, 0, stream>>>(loc, n, n_prot, t_abund);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

size_t csr_workspace_bytes(uint64_t n) { return 16 + ((n + CSR_TILE - 1) / CSR_TILE) * 8 + 16; }

cudaError_t launch_csr(const uint64_t* hash, const uint64_t* loc, uint64_t n, uint64_t* keys, uint32_t* key_grp,
                       uint32_t* grp_start, uint32_t* t_size, uint64_t* d_counts, uint32_t* dir, int dir_bits,
                       int dir_shift, void* workspace, cudaStream_t stream, uint64_t* n_launches) {
    cudaError_t e;
    if (n == 0) {
        e = cudaMemsetAsync(d_counts, 0, 16, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(key_grp, 0, 4, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(grp_start, 0, 4, stream);
        if (e != cudaSuccess) return e;
    } else {
        const uint64_t nt = (n + CSR_TILE - 1) / CSR_TILE;
        e = cudaMemsetAsync(workspace, 0, 16 + nt * 8, stream);
        if (e != cudaSuccess) return e;
        csr_kernel<<<(unsigned)nt, CSR_THREADS, 0, stream>>>(hash, loc, n, keys, key_grp, grp_start, t_size, d_counts,
                                                            (uint32_t*)workspace, (uint64_t*)((char*)workspace + 16));
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_launches) *n_launches += 1;
    }
    const uint64_t work = n ? n : 1;
    unsigned blocks = (unsigned)((work + 255) / 256);
    dir_kernel<<<blocks, 256, 0, stream>>>(keys, d_counts, dir, dir_bits, dir_shift);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

}  // namespace ks
