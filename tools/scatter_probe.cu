// Micro-benchmark: how fast can B200 scatter N 16-byte tuples into 2^tb fixed-capacity buckets with one global
// atomicAdd (with return) per tuple?  Decides whether the sketch kernel can write tuples straight into their sort
// buckets.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_probe scatter_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

template <int MODE>  // 0: atomics only, 1: atomics + 16-byte store, 2: store only (position from hash, no atomic)
__global__ void scatter(uint64_t n, int tb, uint32_t cap, uint32_t* __restrict__ cursor, ulonglong2* __restrict__ out,
                        uint32_t* __restrict__ sink) {
    const uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 8;
    uint32_t acc = 0;
    uint64_t h[8];
    uint32_t pos[8];
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = mix(base + i);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t b = (uint32_t)(h[i] >> (64 - tb));
        if (base + i < n) pos[i] = MODE == 2 ? (uint32_t)(h[i] & 2047u) : atomicAdd(&cursor[b], 1u);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t b = (uint32_t)(h[i] >> (64 - tb));
        if (base + i < n) {
            if (MODE >= 1) { if (pos[i] < cap) out[(uint64_t)b * cap + pos[i]] = make_ulonglong2(h[i], base + i); }
            else acc += pos[i];
        }
    }
    if (MODE == 0 && acc == 0xdeadbeefu) *sink = acc;
}

int main() {
    const uint64_t n = 187000000ull;
    for (int tb = 8; tb <= 17; tb += (tb < 14 ? 3 : 1)) {
        const uint32_t nbk = 1u << tb;
        const uint32_t cap = (uint32_t)(n / nbk * 1.3) + 512;
        uint32_t *cursor, *sink;
        ulonglong2* out;
        cudaMalloc(&cursor, nbk * 4);
        cudaMalloc(&sink, 4);
        cudaMalloc(&out, (uint64_t)nbk * cap * 16);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const unsigned grid = (unsigned)((n / 8 + 255) / 256);
        for (int mode = 0; mode < 3; mode++) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                cudaMemset(cursor, 0, nbk * 4);
                cudaEventRecord(e0);
                if (mode == 0) scatter<0><<<grid, 256>>>(n, tb, cap, cursor, out, sink);
                if (mode == 1) scatter<1><<<grid, 256>>>(n, tb, cap, cursor, out, sink);
                if (mode == 2) scatter<2><<<grid, 256>>>(n, tb, cap, cursor, out, sink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("tb=%2d buckets=%7u cap=%7u mode=%d: %.3f ms  (%.1f G tuples/s)  err=%s\n", tb, nbk, cap, mode, best,
                   n / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(cursor); cudaFree(out); cudaFree(sink);
    }
    return 0;
}
