// In-place lifting form of the time wavelet transform (interleaved ordering),
// /root/reference/source/wavelets.py:81-134: one warp owns the whole time
// column of one space dof in shared memory, runs the J levels there, and
// writes it back: a single HBM pass (8 B read + 8 B written per space-time dof)
// instead of the reference's J sparse products.
#include "stk_common.cuh"

namespace stk {

// a: current values, b: scratch of the same length (N doubles each).
// Level j works on the nodes m*S, m = 0..nj (nj = 2^j, S = 2^(J-j)); even m are
// the hats of level j-1, odd m the wavelets of level j.
__device__ __forceinline__ void synth_level(double *a, double *b, int J, int j, int lane) {
    const int S = 1 << (J - j), nj = 1 << j;
    const double s = exp2(0.5 * j), hs = 0.5 * s;
    for (int m = lane; m <= nj; m += 32) {
        double v;
        if (m & 1) {  // fine[odd] = (c_l + c_r)/2 + s d      (wavelets.py:81-104)
            v = fma(s, a[m * S], 0.5 * (a[(m - 1) * S] + a[(m + 1) * S]));
        } else {  // fine[even] = c - s/2 (d_l + d_r); boundary wavelets count twice
            double dl = a[(m > 0 ? m - 1 : m + 1) * S];
            double dr = a[(m < nj ? m + 1 : m - 1) * S];
            v = fma(-hs, dr, fma(-hs, dl, a[m * S]));
        }
        b[m * S] = v;
    }
    __syncwarp();
    for (int m = lane; m <= nj; m += 32) a[m * S] = b[m * S];
    __syncwarp();
}

__device__ __forceinline__ void analysis_level(double *a, double *b, int J, int j, int lane) {
    const int S = 1 << (J - j), nj = 1 << j;
    const double s = exp2(0.5 * j);
    for (int m = lane; m <= nj; m += 32) {
        double v;
        if (m & 1) {  // d = s (od - ev_l/2 - ev_r/2), boundary hats count twice
            double el = a[(m - 1) * S], er = a[(m + 1) * S];
            v = a[m * S] - 0.5 * el - 0.5 * er;
            if (m == 1) v -= 0.5 * el;
            if (m == nj - 1) v -= 0.5 * er;
            v *= s;
        } else {  // c = ev + od_l/2 + od_r/2
            v = a[m * S];
            if (m > 0) v = fma(0.5, a[(m - 1) * S], v);
            if (m < nj) v = fma(0.5, a[(m + 1) * S], v);
        }
        b[m * S] = v;
    }
    __syncwarp();
    for (int m = lane; m <= nj; m += 32) a[m * S] = b[m * S];
    __syncwarp();
}

template <bool TRANSPOSE>
__global__ void k_wavelet_lift(int M, int J, double *__restrict__ x, int ld) {
    extern __shared__ double smem[];
    const int N = (1 << J) + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    double *a = smem + (size_t)warp * 2 * N;
    double *b = a + N;
    for (int i = blockIdx.x * wpb + warp; i < M; i += gridDim.x * wpb) {
        double *xi = x + (size_t)i * ld;
        for (int t = lane; t < N; t += 32) a[t] = xi[t];
        __syncwarp();
        if (TRANSPOSE) {
            for (int j = J; j >= 1; --j) analysis_level(a, b, J, j, lane);
        } else {
            for (int j = 1; j <= J; ++j) synth_level(a, b, J, j, lane);
        }
        for (int t = lane; t < N; t += 32) xi[t] = a[t];
        __syncwarp();
    }
}

}  // namespace stk

using namespace stk;

extern "C" int stk_wavelet_lift(int M, int J, int transpose, double *x, int ld, void *stream) {
    if (J < 0 || J > 13) return fail(-1, "stk_wavelet_lift: J out of range [0, 13]");
    const int N = (1 << J) + 1;
    if (ld < N) return fail(-1, "stk_wavelet_lift: pitch smaller than 2^J + 1");
    if (M == 0 || J == 0) return 0;
    size_t per_warp = sizeof(double) * 2 * (size_t)N;
    int wpb = (int)((200 * 1024) / per_warp);
    if (wpb > 8) wpb = 8;
    if (wpb < 1) return fail(-1, "stk_wavelet_lift: time column does not fit shared memory");
    size_t smem = per_warp * wpb;
    int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    int64_t want = ((int64_t)M + wpb - 1) / wpb;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    cudaError_t e;
    if (transpose) {
        e = cudaFuncSetAttribute(k_wavelet_lift<true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check(e, "stk_wavelet_lift: smem attribute");
        k_wavelet_lift<true><<<grid, wpb * 32, smem, as_stream(stream)>>>(M, J, x, ld);
    } else {
        e = cudaFuncSetAttribute(k_wavelet_lift<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check(e, "stk_wavelet_lift: smem attribute");
        k_wavelet_lift<false><<<grid, wpb * 32, smem, as_stream(stream)>>>(M, J, x, ld);
    }
    return check_launch("k_wavelet_lift");
}
