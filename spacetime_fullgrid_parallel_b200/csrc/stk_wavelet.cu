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

// A chain of sparse level steps along the time axis of an EXTENDED column
// (local slices followed by halo slices), one warp per space dof, everything
// in shared memory.  Each level lists the rows it changes (CSR over those rows,
// <= 3 entries each for the wavelet lifting steps); their new values go to a
// scratch array and are written back after the level, so a level reads only
// values of the previous one.  This is the lifting form of the wavelet
// transform on a time slab (wavelets.py:81-134 restricted to the slices a rank
// owns plus the <= 2J-1 remote slices its rows depend on): about
// a third of the multiply-adds of the fused matrix, and the adjoint chain
// delivers the partial sums for the remote slices (yh_out) in the same pass.
// Shared layout: tval[NNZ] | A[RT][n_ext] | per warp scratch[n_ext] | ints.
// A CTA takes tiles of RT (32 when it fits) space dofs: the halo slices are slice-major
// (n_halo x M, as the exchange delivers them), so reading -- and, for the
// adjoint, writing -- them for 32 consecutive dofs at a time makes those
// accesses 256-byte runs instead of one 32-byte sector per value.
__global__ void __launch_bounds__(256)
    k_time_chain(const int CHAIN_RT, int M, int n_loc, int n_halo, int nlev, const int *__restrict__ lev_ptr,
                 const int *__restrict__ trow, const int *__restrict__ tptr,
                 const int *__restrict__ tcol, const double *__restrict__ tval, int R, int NNZ,
                 const double *__restrict__ x, int ldx, const double *__restrict__ xh,
                 double *__restrict__ y, int ldy, double *__restrict__ yh_out) {
    extern __shared__ double smem[];
    const int n_ext = n_loc + n_halo;
    const int pe = n_ext | 1;  // odd row pitch: the tile's rows fall into different banks
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *s_val = smem;
    double *At = s_val + NNZ;                          // [CHAIN_RT][pe]
    double *B = At + (size_t)CHAIN_RT * pe + (size_t)warp * n_ext;
    int *s_lev = reinterpret_cast<int *>(At + (size_t)CHAIN_RT * pe + (size_t)wpb * n_ext);
    int *s_row = s_lev + nlev + 1;
    int *s_ptr = s_row + R;
    int *s_col = s_ptr + R + 1;
    for (int k = threadIdx.x; k < NNZ; k += blockDim.x) {
        s_val[k] = __ldg(tval + k);
        s_col[k] = __ldg(tcol + k);
    }
    for (int k = threadIdx.x; k < R; k += blockDim.x) s_row[k] = __ldg(trow + k);
    for (int k = threadIdx.x; k <= R; k += blockDim.x) s_ptr[k] = __ldg(tptr + k);
    for (int k = threadIdx.x; k <= nlev; k += blockDim.x) s_lev[k] = __ldg(lev_ptr + k);
    const int ntiles = (M + CHAIN_RT - 1) / CHAIN_RT;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i0 = tile * CHAIN_RT;
        const int rt = min(CHAIN_RT, M - i0);
        __syncthreads();  // tables staged / previous tile stored
        for (int r = warp; r < rt; r += wpb) {  // local slices: one contiguous run per dof
            const double *xi = x + (size_t)(i0 + r) * ldx;
            for (int t = lane; t < n_loc; t += 32) At[r * pe + t] = xi[t];
        }
        for (int e = threadIdx.x; e < n_halo * CHAIN_RT; e += blockDim.x) {
            const int h = e / CHAIN_RT, r = e - h * CHAIN_RT;
            if (r < rt)
                At[r * pe + n_loc + h] = xh ? __ldg(xh + (size_t)h * M + i0 + r) : 0.0;
        }
        __syncthreads();
        for (int r = warp; r < rt; r += wpb) {
            double *A = At + r * pe;
            for (int lev = 0; lev < nlev; ++lev) {
                const int q0 = s_lev[lev], q1 = s_lev[lev + 1];
                for (int q = q0 + lane; q < q1; q += 32) {
                    double acc = 0.0;
                    for (int p = s_ptr[q]; p < s_ptr[q + 1]; ++p)
                        acc = fma(s_val[p], A[s_col[p]], acc);
                    B[q - q0] = acc;
                }
                __syncwarp();
                for (int q = q0 + lane; q < q1; q += 32) A[s_row[q]] = B[q - q0];
                __syncwarp();
            }
            double *yi = y + (size_t)(i0 + r) * ldy;
            for (int t = lane; t < ldy; t += 32) yi[t] = t < n_loc ? A[t] : 0.0;
        }
        if (yh_out) {
            __syncthreads();
            for (int e = threadIdx.x; e < n_halo * CHAIN_RT; e += blockDim.x) {
                const int h = e / CHAIN_RT, r = e - h * CHAIN_RT;
                if (r < rt) yh_out[(size_t)h * M + i0 + r] = At[r * pe + n_loc + h];
            }
        }
    }
}

}  // namespace stk

using namespace stk;

extern "C" int stk_time_chain(int M, int n_loc, int n_halo, int nlev, const int *lev_ptr,
                              const int *trow, const int *tptr, const int *tcol,
                              const double *tval, int R, int NNZ, const double *x, int ldx,
                              const double *xh, double *y, int ldy, double *yh_out,
                              void *stream) {
    if (x == y) return fail(-1, "stk_time_chain: x must not alias y");
    if (n_loc > ldy || n_loc > ldx) return fail(-1, "stk_time_chain: n_loc exceeds a pitch");
    if (M == 0) return 0;
    const int n_ext = n_loc + n_halo;
    const int wpb = 8;
    const size_t mat = (size_t)NNZ * 12 + (size_t)(2 * R + nlev + 2) * 4 + 16;
    int CHAIN_RT = 32;  // dofs per tile: as many as fit (long time slabs: fewer)
    auto need = [&](int rt) {
        return mat + sizeof(double) * ((size_t)rt * (n_ext | 1) + (size_t)wpb * n_ext);
    };
    while (CHAIN_RT > 1 && need(CHAIN_RT) > 200 * 1024) CHAIN_RT >>= 1;
    const size_t smem = need(CHAIN_RT);
    if (smem > 200 * 1024) return fail(-1, "stk_time_chain: chain does not fit shared memory");
    int ctas = (int)((220 * 1024) / (smem + 1024));
    if (ctas > 6) ctas = 6;
    if (ctas < 1) ctas = 1;
    cudaError_t e = cudaFuncSetAttribute(k_time_chain, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return check(e, "stk_time_chain: smem attribute");
    int64_t want = ((int64_t)M + CHAIN_RT - 1) / CHAIN_RT;
    int64_t cap = (int64_t)sm_count() * ctas;
    k_time_chain<<<(unsigned)(want < cap ? want : cap), wpb * 32, smem, as_stream(stream)>>>(
        CHAIN_RT, M, n_loc, n_halo, nlev, lev_ptr, trow, tptr, tcol, tval, R, NNZ, x, ldx, xh, y, ldy,
        yh_out);
    return check_launch("k_time_chain");
}

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
