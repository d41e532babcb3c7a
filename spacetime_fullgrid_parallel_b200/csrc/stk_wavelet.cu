// In-place lifting form of the time wavelet transform (interleaved ordering),
// /root/reference/source/wavelets.py:81-134: one warp owns the whole time
// column of one space dof in shared memory, runs the J levels there, and
// writes it back: a single HBM pass (8 B read + 8 B written per space-time dof)
// instead of the reference's J sparse products.
#include "stk_common.cuh"

namespace stk {

// Shared memory per warp (one time column): raw[N] | b0[H] | b1[H] | outb[N],
// H = 2^(J-1) + 1.  raw is the input column: level j reads its wavelet
// coefficients (odd nodes) from it.  The hats a level produces are stored
// COMPACTLY (node m of level j at index m) in b[j & 1], so consecutive lanes
// touch consecutive words (no bank conflicts, no shifts) and no level copies
// anything back; the last synthesis level writes straight to global memory,
// the analysis collects finished coefficients in outb for one coalesced store.
template <bool TRANSPOSE>
__global__ void k_wavelet_lift(int M, int J, const double *src, double *x, int ld) {
    extern __shared__ double smem[];
    const int N = (1 << J) + 1, H = (1 << (J - 1)) + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int per_warp = N + 2 * H + (TRANSPOSE ? N : 0);
    double *raw = smem + (size_t)warp * per_warp;
    double *b0 = raw + N, *b1 = b0 + H;
    double *outb = b1 + H;
    for (int i = blockIdx.x * wpb + warp; i < M; i += gridDim.x * wpb) {
        double *xi = x + (size_t)i * ld;
        const double *si = src + (size_t)i * ld;
        for (int t = lane; t < N; t += 32) raw[t] = si[t];
        if (src != x)  // out of place: the pad columns of the result are zero
            for (int t = N + lane; t < ld; t += 32) xi[t] = 0.0;
        __syncwarp();
        if (!TRANSPOSE) {
            // level j: hats c of level j-1 (compact, from the previous level;
            // the two level-0 hats are raw[0], raw[N-1]) and wavelets d of
            // level j (raw, stride 2^(J-j)) -> hats of level j
            // (wavelets.py:81-104)
            for (int j = 1; j <= J; ++j) {
                const int sh = J - j, nj = 1 << j;
                const double s = exp2(0.5 * j), hs = 0.5 * s;
                const double *c = (j - 1) & 1 ? b1 : b0;
                double *dst = j & 1 ? b1 : b0;
                for (int m = lane; m <= nj; m += 32) {
                    double v;
                    if (m & 1) {
                        double cl = (j == 1) ? raw[0] : c[(m - 1) >> 1];
                        double cr = (j == 1) ? raw[N - 1] : c[(m + 1) >> 1];
                        v = fma(s, raw[m << sh], 0.5 * (cl + cr));
                    } else {
                        double cm = (j == 1) ? raw[m ? N - 1 : 0] : c[m >> 1];
                        double dl = raw[(m > 0 ? m - 1 : m + 1) << sh];
                        double dr = raw[(m < nj ? m + 1 : m - 1) << sh];
                        v = fma(-hs, dr, fma(-hs, dl, cm));
                    }
                    if (j == J) xi[m] = v;
                    else dst[m] = v;
                }
                __syncwarp();
            }
        } else {
            // level j (J down to 1): values y of the level-j nodes (compact; the
            // finest level reads raw) -> hats of level j-1 (compact) and the
            // finished level-j wavelet coefficients (wavelets.py:120-134)
            for (int j = J; j >= 1; --j) {
                const int sh = J - j, nj = 1 << j;
                const double s = exp2(0.5 * j);
                const double *y = (j == J) ? raw : ((j + 1) & 1 ? b1 : b0);
                double *dst = j & 1 ? b1 : b0;
                for (int m = lane; m <= nj; m += 32) {
                    if (m & 1) {
                        double el = y[m - 1], er = y[m + 1];
                        double v = y[m] - 0.5 * el - 0.5 * er;
                        if (m == 1) v -= 0.5 * el;
                        if (m == nj - 1) v -= 0.5 * er;
                        outb[m << sh] = v * s;
                    } else {
                        double v = y[m];
                        if (m > 0) v = fma(0.5, y[m - 1], v);
                        if (m < nj) v = fma(0.5, y[m + 1], v);
                        if (j == 1) outb[m << sh] = v;
                        else dst[m >> 1] = v;
                    }
                }
                __syncwarp();
            }
            for (int t = lane; t < N; t += 32) xi[t] = outb[t];
        }
        __syncwarp();
    }
}

}  // namespace stk

using namespace stk;

extern "C" int stk_wavelet_lift(int M, int J, int transpose, const double *src, double *x,
                                int ld, void *stream) {
    if (J < 0 || J > 13) return fail(-1, "stk_wavelet_lift: J out of range [0, 13]");
    const int N = (1 << J) + 1;
    if (ld < N) return fail(-1, "stk_wavelet_lift: pitch smaller than 2^J + 1");
    if (M == 0) return 0;
    if (J == 0) {  // N = 2: the transform is the identity
        if (src != x)
            return check(cudaMemcpyAsync(x, src, sizeof(double) * (size_t)M * ld,
                                         cudaMemcpyDeviceToDevice, as_stream(stream)),
                         "stk_wavelet_lift: copy");
        return 0;
    }
    const int H = (1 << (J - 1)) + 1;
    size_t per_warp = sizeof(double) * ((size_t)N + 2 * H + (transpose ? N : 0));
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
        k_wavelet_lift<true><<<grid, wpb * 32, smem, as_stream(stream)>>>(M, J, src, x, ld);
    } else {
        e = cudaFuncSetAttribute(k_wavelet_lift<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check(e, "stk_wavelet_lift: smem attribute");
        k_wavelet_lift<false><<<grid, wpb * 32, smem, as_stream(stream)>>>(M, J, src, x, ld);
    }
    return check_launch("k_wavelet_lift");
}
