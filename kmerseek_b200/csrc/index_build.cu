// Index build for sm_100a: sort of the sketch tuples by hash and single-pass CSR construction.
//
// Replaces, on the hot path, the reference's per-protein sorted Vec inserts (sourmash KmerMinHash via
// src/rust/signature.rs:273-274), the nested HashMap position records (src/rust/index.rs:770-780) and the
// serial sorted-Vec merge into combined_minhash (src/rust/index.rs:824-827).
#include <cub/device/device_radix_sort.cuh>

#include <vector>

#include "common.cuh"
#include "index_build.cuh"

namespace ks {

namespace {

constexpr int CSR_THREADS = 512;  // big tiles: few tiles in flight keeps the look-back chain short
constexpr int CSR_ROWS = 16;
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

    uint32_t bk[CSR_ROWS], bg[CSR_ROWS];
    uint32_t wk = 0, wg = 0;
#pragma unroll
    for (int r = 0; r < CSR_ROWS; r++) {
        const uint64_t i = i0 + r * 32;
        bool hk = false, hg = false;
        uint64_t h = 0;
        uint32_t pid = 0;
        if (i < n) { h = hash[i]; pid = (uint32_t)(loc[i] >> 32); }
        // the predecessor comes from the neighbouring lane; lane 0 reads it (same cache line as the row before)
        uint64_t hp = __shfl_up_sync(0xffffffffu, h, 1);
        uint32_t pp = __shfl_up_sync(0xffffffffu, pid, 1);
        if (i < n) {
            if (lane == 0 && i > 0) { hp = hash[i - 1]; pp = (uint32_t)(loc[i - 1] >> 32); }
            if (i == 0) {
                hk = hg = true;
            } else {
                hk = h != hp;
                hg = hk || pid != pp;
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
            keys[u] = hash[i];  // L1/L2 hit: this CTA read it a moment ago
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


// ---------------------------------------------------------------------------------------------
// Bucket-local sort (the last pass of the MSD sort).
//
// After the tuples have been partitioned by the top `tb` bits of the normalised hash (normalised =
// shifted left by the `lz` leading bits that are zero under max_hash), every bucket is a contiguous range
// of a few thousand tuples.  One CTA sorts one bucket entirely in shared memory:
//   item = (remaining key bits, left-aligned, low 12 bits replaced by the tuple's index in the bucket)
// so that comparing items compares (hash, original order) exactly -- the sort is stable by construction.
// Two stable 8-bit counting passes (warp-private digit counters, MATCH.ANY ranks) order the items by the
// next 16 key bits; hashes are uniform, so what is left are isolated inversions between neighbours, which
// an odd-even transposition loop removes in a handful of sweeps (it runs until a sweep swaps nothing, so
// the result is exact for any input).  The hash is rebuilt from the item, loc is gathered from a staged
// copy, and the bucket streams back to HBM in sorted order: one read and one write of every tuple.
// ---------------------------------------------------------------------------------------------
constexpr int LS_THREADS = 512;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_CAP = 4096;  // tuples per bucket that fit the shared-memory layout (12 index bits)
constexpr size_t LS_SMEM = (size_t)LS_CAP * 8 * 3 + LS_WARPS * 256 * 2 + 256 * 4 + 64;

__global__ void bucket_start_kernel(const uint64_t* __restrict__ hash, uint64_t n, int lz, int tb,
                                    uint32_t* __restrict__ start, uint32_t* __restrict__ oversize) {
    const uint32_t nb = 1u << tb;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    if (b == 0) oversize[0] = 0;
    if (b == nb) { start[b] = (uint32_t)n; return; }
    uint64_t lo = 0, hi = n;  // first i with bucket(hash[i]) >= b
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint32_t bm = tb ? (uint32_t)((hash[mid] << lz) >> (64 - tb)) : 0u;
        if (bm < b) lo = mid + 1; else hi = mid;
    }
    start[b] = (uint32_t)lo;
}

// Buckets too large for shared memory (only heavy repeats of one hash can do that) are listed for the
// host, which sorts those ranges with the library radix sort.
__global__ void oversize_list_kernel(const uint32_t* __restrict__ start, int tb, uint32_t* __restrict__ oversize,
                                     uint32_t max_list) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (1u << tb)) return;
    if (start[b + 1] - start[b] > (uint32_t)LS_CAP) {
        uint32_t slot = atomicAdd(&oversize[0], 1u);
        if (slot < max_list) oversize[1 + slot] = b;
    }
}

__global__ void __launch_bounds__(LS_THREADS)
bucket_sort_kernel(const uint64_t* __restrict__ in_hash, const uint64_t* __restrict__ in_loc,
                   uint64_t* __restrict__ out_hash, uint64_t* __restrict__ out_loc,
                   const uint32_t* __restrict__ start, int lz, int tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* A = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* B = A + LS_CAP;
    uint64_t* L = B + LS_CAP;
    uint16_t* cnt = reinterpret_cast<uint16_t*>(L + LS_CAP);       // [LS_WARPS][256] warp-private digit counters
    uint32_t* dbase = reinterpret_cast<uint32_t*>(cnt + LS_WARPS * 256);  // [256] exclusive digit offsets
    __shared__ uint32_t s_wsum[8];

    const uint32_t b = blockIdx.x;
    const uint32_t s = start[b], e = start[b + 1];
    const uint32_t m = e - s;
    if (m == 0 || m > (uint32_t)LS_CAP) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sh = lz + tb;  // >= 12 on this path: the low 12 bits of (hash << sh) are zero and carry the index

    // rows of 32 consecutive elements; every warp owns R consecutive rows
    const uint32_t R = (m + LS_THREADS - 1) / LS_THREADS;  // 1..8
    const uint32_t padded = R * LS_THREADS;

    for (uint32_t j = tid; j < padded; j += LS_THREADS) {
        uint64_t item = ~0ull;
        if (j < m) {
            item = ((in_hash[s + j] << sh) & ~0xfffull) | j;
            L[j] = in_loc[s + j];
        }
        A[j] = item;
    }

    uint64_t* src = A;
    uint64_t* dst = B;
    const uint32_t lt = (1u << lane) - 1u;
    uint16_t* my = cnt + warp * 256;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const int dshift = 48 + 8 * pass;
        for (uint32_t i = tid; i < LS_WARPS * 256 / 2; i += LS_THREADS) reinterpret_cast<uint32_t*>(cnt)[i] = 0;
        __syncthreads();
        // sweep 1: lanes of a row that share a digit ("peers") from 8 ballots -- MATCH.ANY would do this in one
        // instruction but runs on the ADU pipe at ~40 cycles per warp; the lowest peer adds the group to the
        // warp-private counter.  digit, rank among peers and group size are kept for sweep 2.
        uint32_t info[8];
#pragma unroll
        for (uint32_t r = 0; r < 8; r++) {
            info[r] = 0;
            if (r < R) {
                const uint32_t d = (uint32_t)(src[(warp * R + r) * 32 + lane] >> dshift) & 255u;
                uint32_t peers = 0xffffffffu;
#pragma unroll
                for (int bit = 0; bit < 8; bit++) {
                    const bool on = (d >> bit) & 1u;
                    const uint32_t v = __ballot_sync(0xffffffffu, on);
                    peers &= on ? v : ~v;
                }
                const uint32_t rank = __popc(peers & lt), size = __popc(peers);
                if (rank == 0) my[d] += (uint16_t)size;
                info[r] = d | (rank << 8) | (size << 16);
                __syncwarp();
            }
        }
        __syncthreads();
        // exclusive scan: over warps within a digit, then over digits
        uint32_t total = 0;
        if (tid < 256) {
#pragma unroll
            for (int w = 0; w < LS_WARPS; w++) {
                uint32_t c = cnt[w * 256 + tid];
                cnt[w * 256 + tid] = (uint16_t)total;
                total += c;
            }
            uint32_t incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            if (lane == 31) s_wsum[warp] = incl;
            total = incl - total;  // exclusive within this warp of digits
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t off = 0;
            for (uint32_t w = 0; w < warp; w++) off += s_wsum[w];
            dbase[tid] = off + total;
        }
        __syncthreads();
        // sweep 2: stable scatter
#pragma unroll
        for (uint32_t r = 0; r < 8; r++) {
            if (r < R) {
                const uint64_t item = src[(warp * R + r) * 32 + lane];
                const uint32_t d = info[r] & 255u, rank = (info[r] >> 8) & 255u, size = info[r] >> 16;
                const uint32_t base = dbase[d] + my[d];
                __syncwarp();
                if (rank == 0) my[d] += (uint16_t)size;
                __syncwarp();
                dst[base + rank] = item;
            }
        }
        __syncthreads();
        uint64_t* t = src; src = dst; dst = t;
    }
    // src == A again.  Remove the remaining inversions (items are distinct, so the order is total).
    while (true) {
        int swapped = 0;
        for (uint32_t i = 2 * tid; i + 1 < padded; i += 2 * LS_THREADS) {
            uint64_t a = src[i], c = src[i + 1];
            if (a > c) { src[i] = c; src[i + 1] = a; swapped = 1; }
        }
        __syncthreads();
        for (uint32_t i = 2 * tid + 1; i + 1 < padded; i += 2 * LS_THREADS) {
            uint64_t a = src[i], c = src[i + 1];
            if (a > c) { src[i] = c; src[i + 1] = a; swapped = 1; }
        }
        if (!__syncthreads_or(swapped)) break;
    }
    const uint64_t top = tb ? ((uint64_t)b << (64 - tb)) : 0ull;
    for (uint32_t j = tid; j < m; j += LS_THREADS) {
        const uint64_t item = src[j];
        out_hash[s + j] = (top | ((item & ~0xfffull) >> tb)) >> lz;
        out_loc[s + j] = L[item & 0xfffu];
    }
}

}  // namespace

size_t sort_temp_bytes(uint64_t n, int end_bit) {
    size_t a = 0, b = 0;
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, v, (int64_t)n, 0, end_bit);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint64_t*)nullptr,
                                    (uint64_t*)nullptr, (int64_t)n, 0, end_bit);
    size_t table = (((size_t)(1u << 24) + 2 + 1024) * 4 + 255) & ~(size_t)255;
    if (n < (1ull << 24)) table = (((size_t)n + 4096 + 2 + 1024) * 4 + 255) & ~(size_t)255;
    return table + (a > b ? a : b) + 256;
}

int msd_top_bits(uint64_t n, int lz) {
    // smallest tb with n / 2^tb <= 3072 (average bucket; LS_CAP = 4096 leaves > 18 sigma for uniform hashes)
    int tb = 0;
    while (tb < 24 && (n >> tb) > 3072) tb++;
    if (lz + tb < 12 || lz + tb > 52) return -1;  // index bits would collide with key bits: library sort instead
    return tb;
}

static cudaError_t library_sort(uint64_t* hash_a, uint64_t* loc_a, uint64_t* hash_b, uint64_t* loc_b, uint64_t n,
                                int begin_bit, int end_bit, void* temp, size_t temp_bytes, cudaStream_t stream,
                                int* out_in_a, uint64_t* n_launches) {
    cub::DoubleBuffer<uint64_t> k(hash_a, hash_b), v(loc_a, loc_b);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, (int64_t)n, begin_bit, end_bit, stream);
    *out_in_a = k.Current() == hash_a ? 1 : 0;
    if (n_launches) *n_launches += 2 + (end_bit - begin_bit + 7) / 8;  // histogram + scan + one onesweep pass per 8 bits
    return e;
}

cudaError_t launch_sort(uint64_t* hash_a, uint64_t* loc_a, uint64_t* hash_b, uint64_t* loc_b, uint64_t n, int end_bit,
                        void* temp, size_t temp_bytes, cudaStream_t stream, int* out_in_a, uint64_t* n_launches) {
    const int lz = 64 - end_bit;
    const int tb = msd_top_bits(n, lz);
    if (tb < 0) return library_sort(hash_a, loc_a, hash_b, loc_b, n, 0, end_bit, temp, temp_bytes, stream, out_in_a, n_launches);

    // 1. partition by the top tb bits (stable): library onesweep passes over those bits only
    int in_a = 1;
    cudaError_t e = cudaSuccess;
    if (tb > 0) {
        e = library_sort(hash_a, loc_a, hash_b, loc_b, n, end_bit - tb, end_bit, temp, temp_bytes, stream, &in_a, n_launches);
        if (e != cudaSuccess) return e;
    }
    uint64_t* sh = in_a ? hash_a : hash_b;
    uint64_t* sl = in_a ? loc_a : loc_b;
    uint64_t* dh = in_a ? hash_b : hash_a;
    uint64_t* dl = in_a ? loc_b : loc_a;
    // 2. bucket boundaries, 3. one CTA per bucket sorts it in shared memory
    const uint32_t nb = 1u << tb;
    uint32_t* start = (uint32_t*)temp;                 // [nb + 1]
    uint32_t* oversize = start + nb + 1;               // [0] = count, then up to MAX_OVERSIZE bucket ids
    constexpr uint32_t MAX_OVERSIZE = 1024;
    bucket_start_kernel<<<(nb + 1 + 255) / 256, 256, 0, stream>>>(sh, n, lz, tb, start, oversize);
    oversize_list_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(start, tb, oversize, MAX_OVERSIZE);
    static bool attr_set = false;
    if (!attr_set) {
        e = cudaFuncSetAttribute(bucket_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    bucket_sort_kernel<<<nb, LS_THREADS, LS_SMEM, stream>>>(sh, sl, dh, dl, start, lz, tb);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_launches) *n_launches += 3;
    // 4. oversize buckets (heavy repeats of few hashes): library sort of the remaining bits, range by range
    uint32_t h_over[1 + MAX_OVERSIZE];
    e = cudaMemcpyAsync(h_over, oversize, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    if (h_over[0]) {
        const uint32_t cnt = h_over[0];
        std::vector<uint32_t> ids;
        std::vector<uint32_t> st(nb + 1);
        e = cudaMemcpy(st.data(), start, (nb + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return e;
        if (cnt <= MAX_OVERSIZE) {
            ids.resize(cnt);
            e = cudaMemcpy(ids.data(), oversize + 1, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) return e;
        } else {
            for (uint32_t b = 0; b < nb; b++) if (st[b + 1] - st[b] > (uint32_t)LS_CAP) ids.push_back(b);
        }
        void* big_temp = (char*)temp + (((size_t)(nb + 2 + MAX_OVERSIZE) * 4 + 255) & ~(size_t)255);
        size_t big_bytes = temp_bytes - ((char*)big_temp - (char*)temp);
        for (uint32_t b : ids) {
            const uint64_t s0 = st[b], m = st[b + 1] - st[b];
            e = cub::DeviceRadixSort::SortPairs(big_temp, big_bytes, sh + s0, dh + s0, sl + s0, dl + s0, (int64_t)m, 0,
                                                end_bit - tb, stream);
            if (e != cudaSuccess) return e;
            if (n_launches) *n_launches += 2 + (end_bit - tb + 7) / 8;
        }
    }
    *out_in_a = in_a ? 0 : 1;
    return cudaSuccess;
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
