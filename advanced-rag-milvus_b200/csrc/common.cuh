// common.cuh -- shared helpers for the b200rag kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/b200rag.h"

namespace b200rag {

// --------------------------------------------------------------------------- error plumbing (api.cu)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();   // bumps the process-wide kernel-launch counter (b200rag_kernel_launch_count)

// A/B options (b200rag_set_option): integer knobs set through an explicit API call and read with relaxed atomic loads;
// -1 = "not set, use the built-in default".  They replace the per-call getenv() lookups of round 1.
enum Option { OPT_SCAN_VERSION = 0, OPT_QG_SPAN, OPT_EPI, OPT_NO_SAMPLE, OPT_STAGE_ROWS, OPT_SAMPLE_MULT, OPT_MMR_PATH,
              OPT_SPARSE_SLICES, OPT_SPARSE_FLAGS, OPT_NO_TIER0, OPT_FINISH_VERSION, OPT_COUNT };
int option(Option o, int dflt);
// Per-thread debug buffers owned by the CALLER (b200rag_debug_set_stats_buffer): device pointer + capacity in u64 slots.
enum StatsKind { STATS_SCAN = 0, STATS_SPARSE = 1, STATS_KINDS };
unsigned long long* stats_buffer(StatsKind k, size_t need_slots);

#define B200_CUDA_CHECK(expr)                                            \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return ::b200rag::cuda_fail(_e, #expr);   \
    } while (0)

#define B200_REQUIRE(cond, ...)                       \
    do {                                              \
        if (!(cond)) {                                \
            ::b200rag::set_error(__VA_ARGS__);        \
            return B200RAG_E_INVALID;                 \
        }                                             \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Workspace {
    char* base;
    size_t cap;
    size_t off = 0;
    Workspace(void* p, size_t n) : base(static_cast<char*>(p)), cap(n) {}
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return p;
    }
    bool ok() const { return off <= cap; }
};

// --------------------------------------------------------------------------- device helpers
// Order-preserving maps float/double -> unsigned (larger value => larger key).
__host__ __device__ __forceinline__ uint32_t mono32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unmono32(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t mono64(double d) {
    uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unmono64(uint64_t k) {
    uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// 16-bit pattern -> fp64 (exact).
template <int DTYPE>
__device__ __forceinline__ double bits_to_double(uint16_t h) {
    if (DTYPE == B200RAG_F16) return (double)__half2float(__ushort_as_half(h));
    return (double)__uint_as_float(((uint32_t)h) << 16);
}
// Two packed 16-bit values -> two doubles (low half first).
template <int DTYPE>
__device__ __forceinline__ void unpack2(uint32_t w, double& a, double& b) {
    if (DTYPE == B200RAG_F16) {
        float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        a = (double)f.x;
        b = (double)f.y;
    } else {
        a = (double)__uint_as_float(w << 16);
        b = (double)__uint_as_float(w & 0xffff0000u);
    }
}

// Canonical dense score of two rows of `dim8` (multiple of 8) 16-bit values.  q given as fp64 (pre-converted),
// x as 16-byte aligned 16-bit row.  8 interleaved fp64 lanes, fixed combine tree (see b200rag.h).
template <int DTYPE>
__device__ __forceinline__ double canonical_dot(const double* __restrict__ q, const uint4* __restrict__ x, int dim8) {
    double p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    for (int c = 0; c < dim8 / 8; ++c) {
        uint4 v = __ldg(x + c);
        double a, b;
        const double* qq = q + c * 8;
        unpack2<DTYPE>(v.x, a, b); p0 = fma(qq[0], a, p0); p1 = fma(qq[1], b, p1);
        unpack2<DTYPE>(v.y, a, b); p2 = fma(qq[2], a, p2); p3 = fma(qq[3], b, p3);
        unpack2<DTYPE>(v.z, a, b); p4 = fma(qq[4], a, p4); p5 = fma(qq[5], b, p5);
        unpack2<DTYPE>(v.w, a, b); p6 = fma(qq[6], a, p6); p7 = fma(qq[7], b, p7);
    }
    return __dadd_rn(__dadd_rn(__dadd_rn(p0, p1), __dadd_rn(p2, p3)),
                     __dadd_rn(__dadd_rn(p4, p5), __dadd_rn(p6, p7)));
}

}  // namespace b200rag
