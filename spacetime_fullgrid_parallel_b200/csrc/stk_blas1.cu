// BLAS-1, layout conversion and the host boundary of libstk.
// Replaces the numpy ufunc passes of /root/reference/source/mpi_vector.py:84-122,
// the np.dot of :205-210 and the axpy pairs of linalg.py:29-30,39-40.
// All kernels are pure HBM streaming: 16-byte vector loads, grid sized as a
// multiple of the SM count, one pass over each operand.
#include "stk_common.cuh"

namespace stk {
thread_local char g_err[512] = "";
int64_t g_launches = 0;

static int g_sms = 0;
int sm_count() {
    if (g_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

// Streaming grid: enough CTAs to cover n2 double2 items, capped at 8 waves of
// 256-thread CTAs per SM (grid-stride beyond that).
static unsigned stream_grid(int64_t n2) {
    int64_t want = (n2 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8 * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (unsigned)want;
}

__global__ void __launch_bounds__(256) k_axpy(double a, const double *__restrict__ x,
                                              double *__restrict__ y, int64_t n2) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
        double2 xv = ldv2(x + 2 * k), yv = ldv2(y + 2 * k);
        yv.x = fma(a, xv.x, yv.x);
        yv.y = fma(a, xv.y, yv.y);
        stv2(y + 2 * k, yv);
    }
}

__global__ void __launch_bounds__(256) k_scale(double a, double *__restrict__ x, int64_t n2) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
        double2 xv = ldv2(x + 2 * k);
        xv.x *= a;
        xv.y *= a;
        stv2(x + 2 * k, xv);
    }
}

// num != nullptr: the scalar is num[0] / den[0], read on the device (Krylov
// coefficients that never visit the host).
__global__ void __launch_bounds__(256) k_xpay(const double *__restrict__ x, double a,
                                              double *__restrict__ y, int64_t n2,
                                              const double *__restrict__ num,
                                              const double *__restrict__ den) {
    if (num) a = __ldg(num) / __ldg(den);
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
        double2 xv = ldv2(x + 2 * k), yv = ldv2(y + 2 * k);
        yv.x = fma(a, yv.x, xv.x);
        yv.y = fma(a, yv.y, xv.y);
        stv2(y + 2 * k, yv);
    }
}

__global__ void __launch_bounds__(256)
    k_pcg_update(double a, const double *__restrict__ p, const double *__restrict__ t,
                 double *__restrict__ w, double *__restrict__ r, int64_t n2,
                 const double *__restrict__ num, const double *__restrict__ den) {
    if (num) a = __ldg(num) / __ldg(den);
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
        double2 pv = ldv2(p + 2 * k), tv = ldv2(t + 2 * k);
        double2 wv = ldv2(w + 2 * k), rv = ldv2(r + 2 * k);
        wv.x = fma(a, pv.x, wv.x);
        wv.y = fma(a, pv.y, wv.y);
        rv.x = fma(-a, tv.x, rv.x);
        rv.y = fma(-a, tv.y, rv.y);
        stv2(w + 2 * k, wv);
        stv2(r + 2 * k, rv);
    }
}

// Single-pass dot: each CTA reduces its grid-stride share with warp shuffles,
// writes one partial; the CTA that takes the last ticket adds the partials in
// index order (deterministic for a fixed grid) and resets the ticket.
// ws[0] (as unsigned) = ticket, ws[1..] = partials.
__global__ void __launch_bounds__(256)
    k_dot(const double *__restrict__ x, const double *__restrict__ y, int64_t n2,
          double *__restrict__ ws, double *__restrict__ out) {
    double acc = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
        double2 xv = ldv2(x + 2 * k), yv = ldv2(y + 2 * k);
        acc = fma(xv.x, yv.x, acc);
        acc = fma(xv.y, yv.y, acc);
    }
    __shared__ double s_part[8];
    __shared__ bool s_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) b += s_part[w];
        ws[1 + blockIdx.x] = b;
        __threadfence();
        unsigned ticket = atomicAdd(reinterpret_cast<unsigned *>(ws), 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // 256 threads add gridDim.x partials: fixed assignment, fixed order.
        double v = 0.0;
        for (unsigned k = threadIdx.x; k < gridDim.x; k += 256)
            v += *((volatile double *)&ws[1 + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) b += s_part[w];
            out[0] = b;
            *reinterpret_cast<unsigned *>(ws) = 0u;
        }
    }
}

// (n_t, M) row-major <-> (M, ld) time-fastest, through a 32x33 shared tile so
// that both the reads and the writes are coalesced.
__global__ void __launch_bounds__(256)
    k_from_rowmajor(const double *__restrict__ src, int n_t, int M, double *__restrict__ dst,
                    int ld) {
    __shared__ double tile[32][33];
    int i0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int t = t0 + r, i = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (t < n_t && i < M) ? src[(size_t)t * M + i] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int i = i0 + r, t = t0 + threadIdx.x;
        if (i < M && t < ld) dst[(size_t)i * ld + t] = tile[threadIdx.x][r];
    }
}

__global__ void __launch_bounds__(256)
    k_to_rowmajor(const double *__restrict__ src, int ld, int n_t, int M,
                  double *__restrict__ dst) {
    __shared__ double tile[32][33];
    int i0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int i = i0 + r, t = t0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < M && t < n_t) ? src[(size_t)i * ld + t] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int t = t0 + r, i = i0 + threadIdx.x;
        if (t < n_t && i < M) dst[(size_t)t * M + i] = tile[threadIdx.x][r];
    }
}

}  // namespace stk

using namespace stk;

extern "C" {

int stk_version(void) { return STK_VERSION; }
const char *stk_last_error(void) { return g_err; }
int64_t stk_launch_count(void) { return g_launches; }
int stk_sync(void *stream) {
    return check(cudaStreamSynchronize(as_stream(stream)), "stk_sync");
}

int stk_axpy(double a, const double *x, double *y, int64_t n, void *stream) {
    if (n & 1) return fail(-1, "stk_axpy: n must be even (pitch is a multiple of 4)");
    if (n == 0) return 0;
    k_axpy<<<stream_grid(n / 2), 256, 0, as_stream(stream)>>>(a, x, y, n / 2);
    return check_launch("k_axpy");
}

int stk_scale(double a, double *x, int64_t n, void *stream) {
    if (n & 1) return fail(-1, "stk_scale: n must be even");
    if (n == 0) return 0;
    k_scale<<<stream_grid(n / 2), 256, 0, as_stream(stream)>>>(a, x, n / 2);
    return check_launch("k_scale");
}

int stk_xpay(const double *x, double a, double *y, int64_t n, void *stream) {
    if (n & 1) return fail(-1, "stk_xpay: n must be even");
    if (n == 0) return 0;
    k_xpay<<<stream_grid(n / 2), 256, 0, as_stream(stream)>>>(x, a, y, n / 2, nullptr, nullptr);
    return check_launch("k_xpay");
}

int stk_xpay_dev(const double *x, const double *num, const double *den, double *y, int64_t n,
                 void *stream) {
    if (n & 1) return fail(-1, "stk_xpay_dev: n must be even");
    if (!num || !den) return fail(-1, "stk_xpay_dev: null scalar");
    if (n == 0) return 0;
    k_xpay<<<stream_grid(n / 2), 256, 0, as_stream(stream)>>>(x, 0.0, y, n / 2, num, den);
    return check_launch("k_xpay");
}

int stk_pcg_update(double a, const double *p, const double *t, double *w, double *r,
                   int64_t n, void *stream) {
    if (n & 1) return fail(-1, "stk_pcg_update: n must be even");
    if (n == 0) return 0;
    k_pcg_update<<<stream_grid(n / 2), 256, 0, as_stream(stream)>>>(a, p, t, w, r, n / 2,
                                                                    nullptr, nullptr);
    return check_launch("k_pcg_update");
}

int stk_pcg_update_dev(const double *num, const double *den, const double *p, const double *t,
                       double *w, double *r, int64_t n, void *stream) {
    if (n & 1) return fail(-1, "stk_pcg_update_dev: n must be even");
    if (!num || !den) return fail(-1, "stk_pcg_update_dev: null scalar");
    if (n == 0) return 0;
    k_pcg_update<<<stream_grid(n / 2), 256, 0, as_stream(stream)>>>(0.0, p, t, w, r, n / 2, num,
                                                                    den);
    return check_launch("k_pcg_update");
}

int stk_dot(const double *x, const double *y, int64_t n, double *ws, double *out_dev,
            void *stream) {
    if (n & 1) return fail(-1, "stk_dot: n must be even");
    int64_t n2 = n / 2;
    int64_t want = (n2 + 256 * 8 - 1) / (256 * 8);  // >= 8 items per thread
    int64_t cap = (int64_t)sm_count() * 8;
    if (cap > STK_DOT_WS - 1) cap = STK_DOT_WS - 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    k_dot<<<(unsigned)want, 256, 0, as_stream(stream)>>>(x, y, n2, ws, out_dev);
    return check_launch("k_dot");
}

int stk_block_from_rowmajor(const double *src, int n_t, int M, double *dst, int ld,
                            void *stream) {
    if (ld < n_t || (ld & 3)) return fail(-1, "stk_block_from_rowmajor: bad pitch");
    if (M == 0) return 0;
    dim3 grid((M + 31) / 32, (ld + 31) / 32), block(32, 8);
    k_from_rowmajor<<<grid, block, 0, as_stream(stream)>>>(src, n_t, M, dst, ld);
    return check_launch("k_from_rowmajor");
}

int stk_block_to_rowmajor(const double *src, int ld, int n_t, int M, double *dst,
                          void *stream) {
    if (ld < n_t) return fail(-1, "stk_block_to_rowmajor: bad pitch");
    if (M == 0 || n_t == 0) return 0;
    dim3 grid((M + 31) / 32, (n_t + 31) / 32), block(32, 8);
    k_to_rowmajor<<<grid, block, 0, as_stream(stream)>>>(src, ld, n_t, M, dst);
    return check_launch("k_to_rowmajor");
}

int stk_block_upload_host(const double *host_rowmajor, int n_t, int M, double *dst, int ld,
                          double *staging, void *stream) {
    cudaStream_t s = as_stream(stream);
    STK_TRY(check(cudaMemcpyAsync(staging, host_rowmajor, sizeof(double) * (size_t)n_t * M,
                                  cudaMemcpyHostToDevice, s),
                  "stk_block_upload_host: H2D"));
    STK_TRY(stk_block_from_rowmajor(staging, n_t, M, dst, ld, stream));
    return check(cudaStreamSynchronize(s), "stk_block_upload_host: sync");
}

int stk_block_download_host(const double *src, int ld, int n_t, int M, double *host_rowmajor,
                            double *staging, void *stream) {
    cudaStream_t s = as_stream(stream);
    STK_TRY(stk_block_to_rowmajor(src, ld, n_t, M, staging, stream));
    STK_TRY(check(cudaMemcpyAsync(host_rowmajor, staging, sizeof(double) * (size_t)n_t * M,
                                  cudaMemcpyDeviceToHost, s),
                  "stk_block_download_host: D2H"));
    return check(cudaStreamSynchronize(s), "stk_block_download_host: sync");
}

}  // extern "C"
