// Shared helpers for libstk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

__device__ __forceinline__ double2 ldv2(const double *p) {
    return *reinterpret_cast<const double2 *>(p);
}
__device__ __forceinline__ void stv2(double *p, double2 v) {
    *reinterpret_cast<double2 *>(p) = v;
}
__device__ __forceinline__ double2 ldg2(const double *p) {
    return __ldg(reinterpret_cast<const double2 *>(p));
}

}  // namespace stk
