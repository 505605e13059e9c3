// Host-side helpers: error type carried to the C ABI boundary, CUDA checks, stream-ordered device arena.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/kmerseek_b200.h"

namespace ks {

struct KsError {
    ks_status status;
    std::string message;
};

[[noreturn]] inline void fail(ks_status s, const std::string& m) { throw KsError{s, m}; }

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return;
    ks_status s = KS_ERR_CUDA;
    if (e == cudaErrorMemoryAllocation) s = KS_ERR_OUT_OF_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) s = KS_ERR_NO_DEVICE;
    (void)cudaGetLastError();
    fail(s, std::string("CUDA error: ") + cudaGetErrorString(e) + " (" + what + " at " + file + ":" + std::to_string(line) + ")");
}
#define KS_CUDA(x) ::ks::cuda_check((x), #x, __FILE__, __LINE__)

// Stream-ordered allocations (cudaMallocAsync on the handle's stream; the pool keeps freed memory, so
// steady-state steps do not hit the driver).  Everything still owned at destruction is freed.
class Arena {
  public:
    Arena(cudaStream_t st, uint64_t* live_bytes) : st_(st), live_(live_bytes) {}
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;
    ~Arena() {
        for (auto& b : blocks_) {
            if (b.p) { cudaFreeAsync(b.p, st_); if (live_) *live_ -= b.bytes; }
        }
    }
    template <class T>
    T* alloc(size_t n) {
        size_t bytes = (n ? n : 1) * sizeof(T);
        bytes = (bytes + 255) & ~(size_t)255;
        void* p = nullptr;
        KS_CUDA(cudaMallocAsync(&p, bytes, st_));
        blocks_.push_back({p, bytes});
        if (live_) *live_ += bytes;
        return (T*)p;
    }
    void release(void* p) {
        for (auto& b : blocks_)
            if (b.p == p && p) { cudaFreeAsync(p, st_); if (live_) *live_ -= b.bytes; b.p = nullptr; return; }
    }
    // Hand ownership of p to `other` (same stream).
    void move_to(Arena& other, void* p) {
        for (auto& b : blocks_)
            if (b.p == p && p) { other.blocks_.push_back(b); b.p = nullptr; return; }
    }
    cudaStream_t stream() const { return st_; }

  private:
    struct Block { void* p; size_t bytes; };
    cudaStream_t st_;
    uint64_t* live_;
    std::vector<Block> blocks_;
};

}  // namespace ks
