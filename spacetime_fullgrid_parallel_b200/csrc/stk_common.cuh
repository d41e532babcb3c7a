// Shared helpers for libstk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <unordered_map>

#include "stk.h"

namespace stk {

extern thread_local char g_err[512];
extern int64_t g_launches;

inline int fail(int code, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

inline int check_launch(const char *kernel) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", kernel, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

inline int check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define STK_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != 0) return _rc;  \
    } while (0)

// The streaming kernels index (row, column-pair) items with 32-bit counters in
// grid-stride loops: `k += stride` must not wrap, so the item count stays one
// full grid (at most 2^21 threads) below 2^32.
#define STK_MAX_ITEMS ((1ll << 32) - (1ll << 21))

inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }

inline unsigned blocks_for(int64_t work, int threads) {
    return (unsigned)((work + threads - 1) / threads);
}

int sm_count();

// Grid for a grid-stride kernel: enough CTAs for `work` items but never more
// than are resident at once (SMs x occupancy), so that no CTA is launched
// behind another one: on B200 a CTA that lives only a few microseconds costs
// ~20 % of the achievable HBM bandwidth in launch/drain overhead (measured,
// profiles/r1_experiments.md).
template <class Kernel>
inline unsigned resident_grid(Kernel kernel, int threads, int64_t work) {
    static thread_local std::unordered_map<const void *, int> cache;
    const void *fn = (const void *)kernel;
    auto it = cache.find(fn);
    if (it == cache.end()) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess ||
            nb < 1)
            nb = 1;
        it = cache.emplace(fn, nb).first;
    }
    int64_t cap = (int64_t)sm_count() * it->second;
    int64_t want = (work + threads - 1) / threads;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

__device__ __forceinline__ double2 ldv2(const double *p) {
    return *reinterpret_cast<const double2 *>(p);
}
__device__ __forceinline__ void stv2(double *p, double2 v) {
    *reinterpret_cast<double2 *>(p) = v;
}
// Four consecutive time values: one 256-bit global access (LDG.E.256 on
// sm_100a; needs 32-byte alignment, which the pitch rule ld % 4 == 0 gives).
struct __align__(32) double4v {
    double x, y, z, w;
};
__device__ __forceinline__ double4v ldv4(const double *p) {
    double4v v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void stv4(double *p, double4v v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z),
                 "d"(v.w)
                 : "memory");
}
__device__ __forceinline__ void fma4(double a, const double4v &x, double4v &s) {
    s.x = fma(a, x.x, s.x);
    s.y = fma(a, x.y, s.y);
    s.z = fma(a, x.z, s.z);
    s.w = fma(a, x.w, s.w);
}

__device__ __forceinline__ double2 ldg2(const double *p) {
    return __ldg(reinterpret_cast<const double2 *>(p));
}

// Row product of a (K-valued) CSR row with the block, for one double2 column:
//   s0 += sum_p v0[p] * x[indices[p], c..c+1]   (s1 likewise with v1, K == 2).
// A plain dependent loop on purpose: fetching 2-8 nonzeros per trip (more
// loads in flight per thread) costs registers, and these kernels are bound by
// occupancy x memory latency -- every batched variant measured slower
// (profiles/r1_experiments.md).
template <int K>
__device__ __forceinline__ void row_product(int p0, int p1, const int *__restrict__ indices,
                                            const double *__restrict__ v0,
                                            const double *__restrict__ v1, const double *x,
                                            int ld, unsigned c, double2 &s0, double2 &s1) {
    for (int p = p0; p < p1; ++p) {
        int j = __ldg(indices + p);
        double2 xv = ldv2(x + (size_t)j * ld + c);
        double a0 = __ldg(v0 + p);
        s0.x = fma(a0, xv.x, s0.x);
        s0.y = fma(a0, xv.y, s0.y);
        if (K == 2) {
            double a1 = __ldg(v1 + p);
            s1.x = fma(a1, xv.x, s1.x);
            s1.y = fma(a1, xv.y, s1.y);
        }
    }
}

}  // namespace stk
