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

inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }

inline unsigned blocks_for(int64_t work, int threads) {
    return (unsigned)((work + threads - 1) / threads);
}

int sm_count();

// Grid for a grid-stride kernel: enough CTAs for `work` items but never more
// than are resident at once (SMs x occupancy), so that no CTA is launched
// behind another one: on B200 a CTA that lives only a few microseconds costs
// ~20 % of the achievable HBM bandwidth in launch/drain overhead (measured,
// profiles/r1b_gs_variants.md).
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
__device__ __forceinline__ double2 ldg2(const double *p) {
    return __ldg(reinterpret_cast<const double2 *>(p));
}

// Row product of a (K-valued) CSR row with the block, for one double2 column:
//   s0 += sum_p v0[p] * x[indices[p], c..c+1]   (s1 likewise with v1, K == 2).
// The first UNROLL nonzeros are loaded as one batch -- all indices and values,
// then all x values -- so that up to UNROLL independent 16-byte loads are in
// flight per thread instead of one dependent chain per nonzero (the rows of
// P1 matrices have 5-9 nonzeros).
template <int K, int UNROLL = 8>
__device__ __forceinline__ void row_product(int p0, int p1, const int *__restrict__ indices,
                                            const double *__restrict__ v0,
                                            const double *__restrict__ v1, const double *x,
                                            int ld, unsigned c, double2 &s0, double2 &s1) {
    if (UNROLL == 0) {  // plain dependent loop
        for (int p = p0; p < p1; ++p) {
            int jj = __ldg(indices + p);
            double2 xx = ldv2(x + (size_t)jj * ld + c);
            double b0 = __ldg(v0 + p);
            s0.x = fma(b0, xx.x, s0.x);
            s0.y = fma(b0, xx.y, s0.y);
            if (K == 2) {
                double b1 = __ldg(v1 + p);
                s1.x = fma(b1, xx.x, s1.x);
                s1.y = fma(b1, xx.y, s1.y);
            }
        }
        return;
    }
    constexpr int U = UNROLL > 0 ? UNROLL : 1;
    int j[U];
    double a0[U], a1[U];
    double2 xv[U];
    const int len = p1 - p0;
#pragma unroll
    for (int q = 0; q < UNROLL; ++q) {
        if (q < len) {
            j[q] = __ldg(indices + p0 + q);
            a0[q] = __ldg(v0 + p0 + q);
            if (K == 2) a1[q] = __ldg(v1 + p0 + q);
        }
    }
#pragma unroll
    for (int q = 0; q < UNROLL; ++q)
        if (q < len) xv[q] = ldv2(x + (size_t)j[q] * ld + c);
#pragma unroll
    for (int q = 0; q < UNROLL; ++q) {
        if (q < len) {
            s0.x = fma(a0[q], xv[q].x, s0.x);
            s0.y = fma(a0[q], xv[q].y, s0.y);
            if (K == 2) {
                s1.x = fma(a1[q], xv[q].x, s1.x);
                s1.y = fma(a1[q], xv[q].y, s1.y);
            }
        }
    }
    for (int p = p0 + UNROLL; p < p1; ++p) {
        int jj = __ldg(indices + p);
        double2 xx = ldv2(x + (size_t)jj * ld + c);
        double b0 = __ldg(v0 + p);
        s0.x = fma(b0, xx.x, s0.x);
        s0.y = fma(b0, xx.y, s0.y);
        if (K == 2) {
            double b1 = __ldg(v1 + p);
            s1.x = fma(b1, xx.x, s1.x);
            s1.y = fma(b1, xx.y, s1.y);
        }
    }
}

}  // namespace stk
