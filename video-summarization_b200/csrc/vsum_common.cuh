// Shared host/device helpers for libvsum_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <atomic>
#include <cstdint>
#include <cstdio>

#include "vsum_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL) && (__CUDA_ARCH__ != 1000)
#error "libvsum_b200 is written for sm_100a only"
#endif

namespace vsum {

// Thread-local error text returned by vsum_last_error().
char *error_buffer();
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);
int eval_sm_budget();       // 0 = unlimited; else the evaluation kernels keep at most this many SMs busy
int scorer_sm_reserve();    // persistent scorer GEMMs leave this many SMs free
int *sched_slot();                // 16 zeroed ints of device memory for one launch's dynamic tile scheduler (nullptr: allocation failed)
int ffn_kernel_version();         // 2 = fused fc1 + ReLU + fc2 + residual + LayerNorm kernel (default), 1 = two GEMM launches
int attention_kernel_version();   // 2 = persistent two-query-tile forward kernel (default), 1 = one 128-query tile per CTA

#define VSUM_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::vsum::set_error(VSUM_ECUDA, "%s failed: %s (%s:%d)", #expr,                \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define VSUM_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                        \
        if (!(cond)) return ::vsum::set_error((code), __VA_ARGS__);                             \
    } while (0)

// Launch check: catches configuration errors at the call site without synchronising.
#define VSUM_LAUNCH_OK(name)                                                                    \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            return ::vsum::set_error(VSUM_ECUDA, "launch of %s failed: %s", name,               \
                                     cudaGetErrorString(_e));                                   \
        ::vsum::count_launch();                                                                 \
    } while (0)

// cudaFuncSetAttribute applies to the CURRENT device only: run `stmt` once per device (a process may drive several
// GPUs through one copy of the library).  Racing threads at worst set the same attribute twice.
#define VSUM_ONCE_PER_DEVICE(stmt)                                                               \
    do {                                                                                        \
        static std::atomic<unsigned long long> _done{0};                                        \
        int _dev = 0;                                                                           \
        VSUM_CUDA_OK(cudaGetDevice(&_dev));                                                     \
        const unsigned long long _bit = 1ull << (_dev & 63);                                    \
        if (!(_done.load(std::memory_order_acquire) & _bit)) {                                  \
            stmt;                                                                               \
            _done.fetch_or(_bit, std::memory_order_release);                                    \
        }                                                                                       \
    } while (0)

// Optional per-kernel timing with CUDA events on the launching stream (vsum_profile_begin/end).
enum ProfCategory {
    PROF_EMBED = 0, PROF_QKV, PROF_ATTN, PROF_OPROJ_LN, PROF_FC1, PROF_FC2_LN, PROF_FFN, PROF_SHOT_MEAN, PROF_KNAPSACK,
    PROF_MASK, PROF_OVERLAP, PROF_FSCORE, PROF_OTHER, PROF_NUM
};
struct ProfScope {
    int idx;
    cudaStream_t stream;
    ProfScope(int category, cudaStream_t s);
    ~ProfScope();
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves 1024-byte aligned sub-buffers out of one caller-provided allocation (base == nullptr: size only).
struct Carver {
    uint8_t *base;
    size_t off = 0;
    template <class T>
    T *get(size_t n) {
        off = align_up(off, 1024);
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

// Largest v in [0, n) with cu[v] <= x (cu ascending, cu[0] = 0, x < cu[n]).
__device__ __forceinline__ int find_segment(const int32_t *__restrict__ cu, int n, int x) {
    int lo = 0, hi = n;            // invariant: cu[lo] <= x < cu[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(cu + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace vsum
